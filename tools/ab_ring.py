"""Developer probe: pipelined step time of the bench workload for the package under <root> (A/B two builds
inside one GPU session).   python tools/ab_ring.py <root> [label]"""
import os
import sys

root = os.path.abspath(sys.argv[1])
sys.path.insert(0, root)
import torch  # noqa: E402
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

N, NB = 131072, 8
W = sampler.sample_rect_worlds(N, NB, seed=1, **sampler.scaled64_params(NB))
wb = WorldBatch(N, NB, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
wb.set_walls(W["wall"])
wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * NB, [(0, (1, 0, 0, 1))] * 4, int(W["rad"].min()), int(W["rad"].max()))
frames = torch.empty((2, N, 64, 64, 4), dtype=torch.uint8, device="cuda")
wb.step_ring_async(20, frames, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = []
for _ in range(4):
    torch.cuda.synchronize()
    e0.record()
    wb.step_ring_async(100, frames, 1)
    e1.record()
    torch.cuda.synchronize()
    out.append(1e3 * e0.elapsed_time(e1) / 100)
torch.cuda.synchronize()
e0.record()
wb.step_async(500)
e1.record()
torch.cuda.synchronize()
print("%-8s ring: %s us/step   physics only: %.1f us/step" % (sys.argv[2] if len(sys.argv) > 2 else root, " ".join("%.1f" % v for v in out),
                                                           1e3 * e0.elapsed_time(e1) / 500))
