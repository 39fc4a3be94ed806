"""BASELINE.json configs[2] and configs[4]: physics-only throughput and per-kernel times for 4..32
balls per world on rectangular and rhombus (polygonal-wall) tables, fp32, one GPU.

    python tools/sweep.py [--out gpurun_out/sweep.json]

Both table kinds come from the device sampler (ppb_sample_rect_worlds; rhombus = isRect=False,
wTheta=[30, 60], dataio.py:243-306).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402


def run(table, n_worlds, n_balls, steps=200, warm=100, max_substeps=64):
    wb = WorldBatch(n_worlds, n_balls, 4, dtype="f32", max_substeps=max_substeps)
    if table == "rect":
        arena = {4: 700, 8: 1200, 16: 1600, 32: 2400}[n_balls]
        wl = {4: (300, 550), 8: (700, 1000), 16: (1000, 1400), 32: (1600, 2200)}[n_balls]
        nf = wb.sample_rect_worlds(seed=3, arena=arena, mn_wlen=wl[0], mx_wlen=wl[1])
        assert nf == 0
    else:
        arena = {4: 1600, 8: 2000, 16: 2600, 32: 3600}[n_balls]
        wl = {4: (500, 800), 8: (800, 1100), 16: (1100, 1500), 32: (1700, 2200)}[n_balls]
        nf = wb.sample_rect_worlds(seed=3, arena=arena, mn_wlen=wl[0], mx_wlen=wl[1], w_theta=[30, 60])
        assert nf == 0
    wb.step_async(warm)
    torch.cuda.synchronize()
    c0 = wb.counters()
    wb.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    wb.step_async(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kms, kn = wb.kernel_times()
    c1 = wb.counters()
    status, _ = wb.get_status()
    live = int((status == 0).sum())
    out = dict(table=table, n_worlds=n_worlds, n_balls=n_balls, steps=steps, ms_per_step=ms / steps,
               ball_steps_per_s=live * n_balls * steps / ms * 1e3, integrate_us=1e3 * kms[0] / max(kn[0], 1),
               collide_us=1e3 * kms[1] / max(kn[1], 1),
               collide_world_frac=(c1["collide_world_steps"] - c0["collide_world_steps"]) / (n_worlds * steps),
               events_per_world_step=(c1["events"] - c0["events"]) / (n_worlds * steps), halted=n_worlds - live,
               max_substeps=max_substeps)
    wb.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    a = ap.parse_args()
    res = []
    for table in ("rect", "rhombus"):
        for nb in (4, 8, 16, 32):
            r = run(table, 65536, nb)
            print(json.dumps(r), flush=True)
            res.append(r)
    res.append(run("rect", 1048576, 4))
    print(json.dumps(res[-1]), flush=True)
    with open(a.out, "w") as fh:
        json.dump(res, fh, indent=1)
