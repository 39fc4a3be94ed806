"""The end-to-end loop of bench.py alone (double-buffered ppb_simulate_host_async), with the frames checked against a
plain device copy: how long a step takes and how the download split itself between the copy engine and the host threads.

    python tools/e2e_async_probe.py [steps]          env: PPB_RAW_FRACTION, PPB_UNPACK_THREADS, PPB_NO_AVX512
"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyphy_engine_b200 import sampler, sharding
from pyphy_engine_b200.engine import WorldBatch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n, nb = 131072, 8
W = sharding.sample_world_range(0, n, nb, seed0=1, **sampler.scaled64_params(nb))
wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
traj = [torch.empty((1, n, nb, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
frames = [torch.empty((1, n, 64, 64, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
wb.simulate_host(1, traj[0], frames[0], 1)
wb.simulate_host(1, traj[1], frames[1], 1)

# correctness: the same step through the device-resident path, downloaded raw
wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
dev_frames = torch.empty((1, n, 64, 64, 4), dtype=torch.uint8, device="cuda")
wb.step_async(1, None, dev_frames, 1)
torch.cuda.synchronize()
want = dev_frames.cpu()
wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
frames[0].zero_()
wb.simulate_host(1, traj[0], frames[0], 1)
print("frames identical to the raw copy:", bool(torch.equal(want, frames[0])))

c0 = wb.counters()
per = []
t0 = time.perf_counter()
for i in range(steps):
    ts = time.perf_counter()
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.simulate_host(1, traj[i & 1], frames[i & 1], 1, wait=False)
    wb.host_sync(keep_in_flight=1)
    per.append(1e3 * (time.perf_counter() - ts))
wb.host_sync()
dt = time.perf_counter() - t0
c1 = wb.counters()
print("same after the loop:", bool(torch.equal(want, frames[(steps - 1) & 1])))
d = {k: c1[k] - c0[k] for k in c1 if k.startswith(("d2h", "unpack"))}
print("ms per step %.2f  (per step: %s)" % (1e3 * dt / steps, " ".join("%.1f" % x for x in per)))
print("ball-steps/s %.3e" % (n * nb * steps / dt))
print("per step: packed %.1f MB  raw %.1f MB  unpack %.2f ms" % (d["d2h_packed_bytes"] / steps / 1e6, d["d2h_raw_bytes"] / steps / 1e6,
                                                                  1e-3 * d["unpack_us"] / steps))
