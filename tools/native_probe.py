"""Developer probe: render-only throughput at native canvas sizes (ICLR16 arena 700, r = 25).
    python tools/native_probe.py [worlds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200 import sampler  # noqa: E402
from pyphy_engine_b200.engine import WorldBatch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for arena in (700, 667, 1200):
    nb = 8 if arena == 1200 else 4
    kw = dict(arena=arena, mn_wlen=700, mx_wlen=1000) if arena == 1200 else dict(arena=arena, mn_wlen=300, mx_wlen=550)
    W = sampler.sample_rect_worlds(n, nb, seed=1, **kw)
    wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(arena, arena), max_substeps=64)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.set_styles([(2, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 25, 25)
    frames = torch.empty((n, arena, arena, 4), dtype=torch.uint8, device="cuda")
    wb.step_async(20)
    wb.render(out=frames)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        wb.render(out=frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("arena %4d  %d worlds x %d balls: render %.3f ms  %.3e frames/s  %.2f TB/s" % (
        arena, n, nb, ms, n / ms * 1e3, frames.numel() / ms / 1e9), flush=True)
    wb.close()
    del frames
