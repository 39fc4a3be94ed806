"""Developer probe: time of one advance launch as a function of the steps it covers (physics only), back to back
(the next launch's head overlaps this one's tail) and isolated (a synchronisation after every launch)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.environ.get("PPB_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # PPB_ROOT: another build (A/B)
import torch  # noqa: E402
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

for N, NB in ((131072, 8), (65536, 4)):
    W = sampler.sample_rect_worlds(N, NB, seed=1, **sampler.scaled64_params(NB))
    wb = WorldBatch(N, NB, 4, dtype="f32", max_substeps=64)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.step_async(100)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for n in ((50, 64) if os.environ.get("PROBE_SHORT") else (1, 4, 16, 20, 64)):
        reps = max(4, 128 // n)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            wb.step_async(n)
        e1.record()
        torch.cuda.synchronize()
        t = 1e3 * e0.elapsed_time(e1) / reps
        iso = 0.0
        for _ in range(reps):
            torch.cuda.synchronize()
            e0.record()
            wb.step_async(n)
            e1.record()
            torch.cuda.synchronize()
            iso += 1e3 * e0.elapsed_time(e1) / reps
        print("%d x %d: %2d steps per launch: back to back %.1f us per launch (%.1f per step), isolated %.1f us (%.1f per step)"
              % (N, NB, n, t, t / n, iso, iso / n), flush=True)
    wb.close()
