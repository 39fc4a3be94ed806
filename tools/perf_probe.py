"""Developer probe: per-kernel times of the step pipeline for a few batch shapes.

    python tools/perf_probe.py [--worlds N] [--balls B] [--steps S] [--render]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.environ.get("PPB_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # PPB_ROOT: another build (A/B)
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402


def run(n_worlds, n_balls, steps, render, dtype="f32", px64=True, warm=100, max_substeps=256):
    kw = sampler.scaled64_params(n_balls) if px64 else {}
    t0 = time.time()
    W = sampler.sample_rect_worlds(n_worlds, n_balls, seed=1, **kw)
    t_sample = time.time() - t0
    canvas = (64, 64) if px64 else ((int(os.environ.get("PPB_PROBE_CANVAS", "700")),) * 2)
    wb = WorldBatch(n_worlds, n_balls, 4, dtype=dtype, canvas=canvas, max_substeps=max_substeps)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    st = 1 if px64 else 2
    wb.set_styles([(st, (.5, .5, .5, 1), (0, 0, 0, 1))] * n_balls, [(0, (1, 0, 0, 1))] * 4,
                  int(W["rad"].min()), int(W["rad"].max()))
    frames = None
    if render:
        frames = torch.empty((2, n_worlds, canvas[1], canvas[0], 4), dtype=torch.uint8, device="cuda")
    wb.step_async(warm)
    torch.cuda.synchronize()
    c0 = wb.counters()
    wb.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if render:
        wb.step_ring_async(steps, frames, 1)
    else:
        wb.step_async(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = wb.counters()
    kms, kn = wb.kernel_times()
    status, _ = wb.get_status()
    cw = c1["collide_world_steps"] - c0["collide_world_steps"]
    it = c1["collide_iterations"] - c0["collide_iterations"]
    print("worlds %d balls %d dtype %s render %d: %.3f ms/step  %.3e ball-steps/s  | advance %.1f us/launch (%d launches) render %.1f us/launch | "
          "event-loop world-steps %.1f%%  iters/event-loop step %.2f  events/world-step %.3f  halted %d %s (sample %.1fs)" % (
              n_worlds, n_balls, dtype, render, ms / steps, n_worlds * n_balls * steps / ms * 1e3,
              1e3 * kms[0] / max(kn[0], 1), kn[0], 1e3 * kms[2] / max(kn[2], 1),
              100.0 * cw / (n_worlds * steps), it / max(cw, 1), (c1["events"] - c0["events"]) / (n_worlds * steps),
              int((status != 0).sum()), np.bincount(status, minlength=10).tolist(), t_sample), flush=True)
    wb.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=0)
    ap.add_argument("--balls", type=int, default=4)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--render", action="store_true")
    ap.add_argument("--native", action="store_true")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--warm", type=int, default=100)
    ap.add_argument("--max-substeps", type=int, default=256)
    a = ap.parse_args()
    if a.worlds:
        run(a.worlds, a.balls, a.steps, a.render, a.dtype, not a.native, a.warm, a.max_substeps)
    else:
        run(65536, 4, 200, False, "f32", False, 100, 64)
        run(1048576, 4, 100, False, "f32", True, 100, 64)
        run(131072, 8, 100, False, "f32", True, 100, 64)
        run(131072, 8, 100, False, "f32", True, 100, 16)
        run(131072, 8, 50, True, "f32", True, 100, 64)
        run(65536, 4, 100, False, "f64", False, 100, 4096)
