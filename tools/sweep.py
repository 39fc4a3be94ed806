"""BASELINE.json configs[2] and configs[4]: physics-only throughput and per-kernel times for 4..32
balls per world on rectangular, rhombus and regular-hexagon (polygonal-wall) tables, fp32, one GPU --
or, under torchrun, 65,536 worlds on EVERY GPU (weak scaling; per-rank seeds; no collective in the step loop;
time = max over ranks, ball-steps summed over ranks).

    python tools/sweep.py [--out gpurun_out/sweep.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py --out gpurun_out/sweep_8gpu.json

Rectangles and rhombi come from the device sampler (ppb_sample_rect_worlds; rhombus = isRect=False,
wTheta=[30, 60], dataio.py:243-306); the hexagon is one create_cage table (primitives.py:199-209)
shared by all worlds, with per-world ball places and velocities drawn on the host.

Every shape is run twice: once plain (ms_per_step / ball_steps_per_s: what a caller gets) and once with
the library's per-kernel CUDA events switched on (advance_us_per_launch; the events themselves cost
about as much as the kernels at this size, so that pass is only used for the split).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD_SIZE = int(os.environ.get("WORLD_SIZE", "1"))


def hexagon_worlds(wb, n_worlds, n_balls, seed=3, w_thick=30, r_ball=25):
    """One regular hexagon cage for every world; balls on distinct cells of a lattice inside the
    inscribed circle (so they never overlap), random directions, DataSaver's speed range."""
    from pyphy_engine_b200 import geometry as gm, primitives as pm
    rs = np.random.RandomState(seed)
    spacing = 2 * r_ball + 6
    cells_needed = 12 * n_balls                                       # about the rect tables' free area per ball
    m = int(np.ceil(np.sqrt(cells_needed * 4 / np.pi))) + 1          # lattice cells across the inscribed circle
    free = m * spacing / 2.0
    radius = (free + w_thick + r_ball + 2) / np.cos(np.pi / 6)
    c = np.ceil(radius) + 10
    pts = [gm.Point(c + radius * np.cos(2 * np.pi * i / 6), c + radius * np.sin(2 * np.pi * i / 6)) for i in range(6)]
    wv = np.zeros((1, 6, 4, 2), np.float64)
    for i, w in enumerate(pm.create_cage(pts, wThick=w_thick)):
        for k, l in enumerate(w.get_lines()):
            wv[0, i, k] = (l.st().x(), l.st().y())
    g = (np.arange(m) - (m - 1) / 2.0) * spacing
    gx, gy = np.meshgrid(g, g)
    ok = (gx ** 2 + gy ** 2) <= (free - r_ball) ** 2
    cells = np.stack([gx[ok], gy[ok]], 1) + c
    assert len(cells) >= n_balls, (len(cells), n_balls)
    pick = np.argsort(rs.rand(n_worlds, len(cells)), axis=1)[:, :n_balls]
    pos = np.floor(cells[pick])
    mass = 4.0 / 3.0 * np.pi * (r_ball / 10.0) ** 3
    speed = 0.5 * (3e4 + np.floor(rs.rand(n_worlds, n_balls) * 5e4)) / mass
    ang = 2 * np.pi * rs.rand(n_worlds, n_balls)
    vel = np.stack([speed * np.cos(ang), speed * np.sin(ang)], -1)
    wb.set_walls(np.broadcast_to(wv, (n_worlds, 6, 4, 2)).copy())
    wb.set_balls(pos, vel, np.full((n_worlds, n_balls), float(r_ball)), np.full((n_worlds, n_balls), mass))


def run(table, n_worlds, n_balls, steps=500, warm=100, max_substeps=64, reps=3):
    wb = WorldBatch(n_worlds, n_balls, 6 if table == "hexagon" else 4, dtype="f32", max_substeps=max_substeps,
                    device=LOCAL_RANK)
    seed = 3 + 1000 * RANK
    if table == "hexagon":
        hexagon_worlds(wb, n_worlds, n_balls, seed=seed)
    elif table == "rect":
        arena = {4: 700, 8: 1200, 16: 1600, 32: 2400}[n_balls]
        wl = {4: (300, 550), 8: (700, 1000), 16: (1000, 1400), 32: (1600, 2200)}[n_balls]
        nf = wb.sample_rect_worlds(seed=seed, arena=arena, mn_wlen=wl[0], mx_wlen=wl[1])
        assert nf == 0
    else:
        arena = {4: 1600, 8: 2000, 16: 2600, 32: 3600}[n_balls]
        wl = {4: (500, 800), 8: (800, 1100), 16: (1100, 1500), 32: (1700, 2200)}[n_balls]
        nf = wb.sample_rect_worlds(seed=seed, arena=arena, mn_wlen=wl[0], mx_wlen=wl[1], w_theta=[30, 60])
        assert nf == 0
    wb.step_async(warm)
    torch.cuda.synchronize()
    c0 = wb.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 1e30
    for _ in range(reps):           # best of `reps` back-to-back runs of `steps` steps (max over ranks)
        if WORLD_SIZE > 1:
            torch.distributed.barrier()
        e0.record()
        wb.step_async(steps)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if WORLD_SIZE > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = min(ms, float(t[0]))
    c1 = wb.counters()
    wb.set_profiling(True)          # second pass, only for the per-kernel split
    wb.step_async(steps)
    torch.cuda.synchronize()
    kms, kn = wb.kernel_times()
    wb.set_profiling(False)
    status, _ = wb.get_status()
    live = int((status == 0).sum())
    n_local = n_worlds
    if WORLD_SIZE > 1:
        tot = torch.tensor([live, n_worlds], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(tot)
        live, n_worlds = int(tot[0]), int(tot[1])
    out = dict(table=table, n_gpus=WORLD_SIZE, n_worlds=n_worlds, n_balls=n_balls, steps=steps, ms_per_step=ms / steps,
               ball_steps_per_s=live * n_balls * steps / ms * 1e3, advance_us_per_launch=1e3 * kms[0] / max(kn[0], 1),
               advance_launches=kn[0],
               event_loop_world_frac=(c1["collide_world_steps"] - c0["collide_world_steps"]) / (n_local * steps * reps),
               events_per_world_step=(c1["events"] - c0["events"]) / (n_local * steps * reps), halted=n_worlds - live,
               max_substeps=max_substeps)
    wb.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    a = ap.parse_args()
    if WORLD_SIZE > 1:
        torch.cuda.set_device(LOCAL_RANK)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", LOCAL_RANK))
    res = []
    for table in ("rect", "rhombus", "hexagon"):
        for nb in (4, 8, 16, 32):
            r = run(table, 65536, nb)
            if RANK == 0:
                print(json.dumps(r), flush=True)
            res.append(r)
    res.append(run("rect", 1048576, 4))
    if RANK == 0:
        print(json.dumps(res[-1]), flush=True)
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)
    if WORLD_SIZE > 1:
        torch.distributed.destroy_process_group()
