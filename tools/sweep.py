"""BASELINE.json configs[2] and configs[4]: physics-only throughput and per-kernel times for 4..32
balls per world on rectangular and rhombus (polygonal-wall) tables, fp32, one GPU.

    python tools/sweep.py [--out gpurun_out/sweep.json]

Rectangular tables come from the device sampler; rhombus tables from the stream-exact DataSaver
(isRect=False, dataio.py:243-306) -- 512 sampled tables tiled to the batch size.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402


def rhombus_worlds(n_worlds, n_balls, seed=1, n_src=512):
    import pyphy_engine_b200.dataio as dio
    arena = {4: 1600, 8: 2000, 16: 2600, 32: 3600}[n_balls]
    wl = {4: (500, 800), 8: (800, 1100), 16: (1100, 1500), 32: (1700, 2200)}[n_balls]
    ds = dio.DataSaver(rootPath=tempfile.mkdtemp(), makeDirs=False, randSeed=seed, numBalls=n_balls, mnBallSz=25,
                       mxBallSz=25, mnForce=3e4, mxForce=8e4, wThick=30, isRect=False, wTheta=[30, 60],
                       mnWLen=wl[0], mxWLen=wl[1], arenaSz=arena, mnSeqLen=10, mxSeqLen=10)
    specs = ds.sample_sequences(n_src)
    wall = np.zeros((n_src, 4, 4, 2)); pos = np.zeros((n_src, n_balls, 2)); vel = np.zeros_like(pos)
    for s, sp in enumerate(specs):
        for i, w in enumerate(sp.model.world_.static_.values()):
            wall[s, i] = [[v.x(), v.y()] for v in w.bbox_.vert_]
        for j, b in enumerate(sp.model.world_.dynamic_.values()):
            pos[s, j] = (b.pos_.x(), b.pos_.y()); vel[s, j] = (b.vel_.x(), b.vel_.y())
    rep = -(-n_worlds // n_src)
    t = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:n_worlds]
    m = (4 / 3.0) * np.pi * 2.5 ** 3
    return t(wall), t(pos), t(vel), np.full((n_worlds, n_balls), 25.0), np.full((n_worlds, n_balls), m)


def run(table, n_worlds, n_balls, steps=200, warm=100, max_substeps=64):
    wb = WorldBatch(n_worlds, n_balls, 4, dtype="f32", max_substeps=max_substeps)
    if table == "rect":
        arena = {4: 700, 8: 1200, 16: 1600, 32: 2400}[n_balls]
        wl = {4: (300, 550), 8: (700, 1000), 16: (1000, 1400), 32: (1600, 2200)}[n_balls]
        nf = wb.sample_rect_worlds(seed=3, arena=arena, mn_wlen=wl[0], mx_wlen=wl[1])
        assert nf == 0
    else:
        wall, pos, vel, rad, mass = rhombus_worlds(n_worlds, n_balls)
        wb.set_walls(wall)
        wb.set_balls(pos, vel, rad, mass)
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
