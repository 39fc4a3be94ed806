"""Developer probe: how long are the sub-step chains the collide kernel has to follow?

Runs the physics-only sweep shape (65,536 worlds, device sampler) with the event ring enabled and prints
  * the histogram of events per (world, step) pair -- a world with n events in one step went through
    at least n/2 sub-steps, each with a full rescan of its balls;
  * the collide kernel's time per launch for several sub-step limits (a limit of 2 halts the rare long
    chains, so the difference to the default is what the longest chain of a launch costs).

    python tools/chain_probe.py [--balls 4] [--worlds 65536]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get("PPB_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

ARENA = {4: 700, 8: 1200, 16: 1600, 32: 2400}
WLEN = {4: (300, 550), 8: (700, 1000), 16: (1000, 1400), 32: (1600, 2200)}


def make(n_worlds, n_balls, max_substeps, event_capacity=0):
    wb = WorldBatch(n_worlds, n_balls, 4, dtype="f32", max_substeps=max_substeps, event_capacity=event_capacity)
    nf = wb.sample_rect_worlds(seed=3, arena=ARENA[n_balls], mn_wlen=WLEN[n_balls][0], mx_wlen=WLEN[n_balls][1])
    assert nf == 0
    return wb


def histogram(n_worlds, n_balls, steps=200, warm=100):
    wb = make(n_worlds, n_balls, 64, event_capacity=int(n_worlds * steps * 0.6) + 1024)
    wb.step_async(warm)
    wb.clear_events()
    wb.step_async(steps)
    torch.cuda.synchronize()
    ev, total = wb.read_events()
    assert total <= wb.event_capacity, (total, wb.event_capacity)
    key = ev[:, 0].astype(np.int64) * 100000 + ev[:, 1].astype(np.int64)
    ks, idx, cnt = np.unique(key, return_index=True, return_counts=True)
    h = np.bincount(cnt)
    print("events %d in %d (world, step) pairs; events per pair histogram:" % (len(key), len(cnt)))
    print("  " + "  ".join("%d:%d" % (i, h[i]) for i in range(1, len(h)) if h[i]))
    stp = ev[idx, 1].astype(np.int64)
    stp -= stp.min()
    mx = np.zeros(int(stp.max()) + 1, np.int64)
    np.maximum.at(mx, stp, cnt)
    mx = mx[mx > 0]
    print("  longest chain of a launch (events in one world): median %d  p90 %d  max %d" % (
        np.median(mx), np.percentile(mx, 90), mx.max()))
    wb.close()


def timing(n_worlds, n_balls, limit, steps=200, warm=100):
    wb = make(n_worlds, n_balls, limit)
    wb.step_async(warm)
    torch.cuda.synchronize()
    wb.set_profiling(True)
    c0 = wb.counters()
    wb.step_async(steps)
    torch.cuda.synchronize()
    c1 = wb.counters()
    kms, kn = wb.kernel_times()
    status, _ = wb.get_status()
    cw = c1["collide_world_steps"] - c0["collide_world_steps"]
    print("limit %3d: K1 %.1f us  K2 %.1f us  listed/launch %.0f  iters/listed %.2f  rescans/listed %.2f  halted %d" % (
        limit, 1e3 * kms[0] / kn[0], 1e3 * kms[1] / kn[1], cw / steps,
        (c1["collide_iterations"] - c0["collide_iterations"]) / max(cw, 1),
        (c1["rescans"] - c0["rescans"]) / max(cw, 1), int((status != 0).sum())), flush=True)
    wb.close()


def plain(n_worlds, n_balls, steps=1000, warm=100):
    """ms/step with no profiling events between the kernels (what a user gets)."""
    wb = make(n_worlds, n_balls, 64)
    wb.step_async(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        wb.step_async(steps)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    status, _ = wb.get_status()
    live = int((status == 0).sum())
    print("plain: worlds %d balls %d  %.2f us/step  %.3e ball-steps/s  (PPB_NO_PDL=%s)" % (
        n_worlds, n_balls, 1e3 * best, live * n_balls / best * 1e3, os.environ.get("PPB_NO_PDL", "0")), flush=True)
    wb.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--balls", type=int, default=4)
    ap.add_argument("--worlds", type=int, default=65536)
    ap.add_argument("--plain", action="store_true")
    a = ap.parse_args()
    if a.plain:
        plain(a.worlds, a.balls)
        sys.exit(0)
    histogram(a.worlds, a.balls)
    for limit in (2, 3, 4, 8, 64):
        timing(a.worlds, a.balls, limit)
