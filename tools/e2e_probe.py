"""Where the end-to-end (host buffers) step spends its time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200 import sampler, sharding
from pyphy_engine_b200.engine import WorldBatch
n, nb = 131072, 8
W = sharding.sample_world_range(0, n, nb, seed0=1, **sampler.scaled64_params(nb))
wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
traj = torch.empty((1, n, nb, 2), dtype=torch.float32).pin_memory()
frames = torch.empty((1, n, 64, 64, 4), dtype=torch.uint8).pin_memory()
wb.simulate_host(1, traj, frames, 1)
T = {"set_walls": 0.0, "set_balls": 0.0, "simulate_host": 0.0}
for _ in range(8):
    t0 = time.perf_counter(); wb.set_walls(W["wall"]); t1 = time.perf_counter()
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"]); t2 = time.perf_counter()
    wb.simulate_host(1, traj, frames, 1); t3 = time.perf_counter()
    T["set_walls"] += t1 - t0; T["set_balls"] += t2 - t1; T["simulate_host"] += t3 - t2
print({k: round(1e3 * v / 8, 2) for k, v in T.items()}, "ms per step; D2H", frames.numel() / 1e9, "GB ->", round(frames.numel() / 1e9 / (T["simulate_host"] / 8), 1), "GB/s incl. compute")
print({k: v for k, v in wb.counters().items() if k.startswith(("d2h", "unpack"))})
# raw pinned D2H bandwidth for reference
d = torch.empty(frames.shape, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4): frames.copy_(d, non_blocking=True)
torch.cuda.synchronize(); print("raw D2H GB/s", round(4 * frames.numel() / 1e9 / (time.perf_counter() - t0), 1))
