"""Developer probe: can an advance launch that occupies only a few warps per SM (PPB_ADV_OCC CTAs per SM) run beside the
render launches without slowing them?   PPB_ADV_OCC=1 python tools/overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.environ.get("PPB_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

N, NB = 131072, 8
W = sampler.sample_rect_worlds(N, NB, seed=1, **sampler.scaled64_params(NB))
wb = WorldBatch(N, NB, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
wb.set_walls(W["wall"])
wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * NB, [(0, (1, 0, 0, 1))] * 4, 3, 3)
frames = torch.empty((2, N, 64, 64, 4), dtype=torch.uint8, device="cuda")
wb.step_async(20)
wb.render(out=frames[0])
s2 = torch.cuda.Stream(priority=-1)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
R = int(os.environ.get("PROBE_RENDERS", "24"))
STEPS = int(os.environ.get("PROBE_STEPS", "24"))
for rep in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    wb.step_async(STEPS)
    ev[1].record()
    torch.cuda.synchronize()
    adv_alone = ev[0].elapsed_time(ev[1])
    ev[2].record(s2)
    for _ in range(R):
        wb.render(out=frames[0], stream=s2)
    ev[3].record(s2)
    torch.cuda.synchronize()
    rend_alone = ev[2].elapsed_time(ev[3])
    # both: the advance first (it takes its places), the renders beside it
    ev[0].record()
    wb.step_async(STEPS)
    ev[1].record()
    s2.wait_event(ev[0])
    ev[2].record(s2)
    for _ in range(R):
        wb.render(out=frames[0], stream=s2)
    ev[3].record(s2)
    torch.cuda.synchronize()
    print("advance(%d steps) alone %.0f us | %d renders alone %.0f us (%.1f each) | together: advance %.0f us, renders %.0f us (%.1f each), both done after %.0f us"
          % (STEPS, 1e3 * adv_alone, R, 1e3 * rend_alone, 1e3 * rend_alone / R, 1e3 * ev[0].elapsed_time(ev[1]),
             1e3 * ev[2].elapsed_time(ev[3]), 1e3 * ev[2].elapsed_time(ev[3]) / R, 1e3 * max(ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[3]))), flush=True)
wb.close()
