"""Shard equivalence on real devices: the worlds of a job give identical results whether they run
as one batch on one GPU or as shards on several (no cross-world coupling).
The two-device case needs >= 2 GPUs (``gpurun --gpus 2``); it is skipped on a single-GPU box."""
import numpy as np
import pytest
import torch

from pyphy_engine_b200 import sampler, sharding
from pyphy_engine_b200.engine import WorldBatch

pytestmark = pytest.mark.gpu


def _run(W, device, steps, dtype):
    n, nb = W["pos"].shape[:2]
    wb = WorldBatch(n, nb, 4, dtype=dtype, canvas=(64, 64), device=device, max_substeps=256, event_capacity=1 << 20)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
    with torch.cuda.device(device):
        traj, frames = wb.step(steps, record_traj=True, render_every=steps)
        out = traj.cpu().numpy(), frames.cpu().numpy(), wb.get_status()[0], wb.read_events()[0]
    wb.close()
    return out


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_shards_on_one_device_equal_the_whole_batch(dtype):
    n, nb, steps = 3000, 4, 60
    kw = sampler.scaled64_params(nb)
    full = sharding.sample_world_range(0, n, nb, seed0=5, **kw)
    t0, f0, s0, e0 = _run(full, 0, steps, dtype)
    for g in (2, 3):
        ts, fs, ss, es = [], [], [], []
        for r in range(g):
            first, cnt = sharding.shard_range(n, r, g)
            t, f, s, e = _run(sharding.sample_world_range(first, cnt, nb, seed0=5, **kw), 0, steps, dtype)
            e = e.copy()
            e[:, 0] += first
            ts.append(t); fs.append(f); ss.append(s); es.append(e)
        assert np.array_equal(np.concatenate(ts, 1), t0)
        assert np.array_equal(np.concatenate(fs, 1), f0)
        assert np.array_equal(np.concatenate(ss), s0)
        assert np.array_equal(np.concatenate(es), e0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_equal_one():
    n, nb, steps = 8192, 8, 40
    kw = sampler.scaled64_params(nb)
    full = sharding.sample_world_range(0, n, nb, seed0=9, **kw)
    t0, f0, s0, e0 = _run(full, 0, steps, "f32")
    parts = []
    for r in range(2):
        first, cnt = sharding.shard_range(n, r, 2)
        parts.append(_run(sharding.sample_world_range(first, cnt, nb, seed0=9, **kw), r, steps, "f32"))
    assert np.array_equal(np.concatenate([p[0] for p in parts], 1), t0)
    assert np.array_equal(np.concatenate([p[1] for p in parts], 1), f0)
    assert np.array_equal(np.concatenate([p[2] for p in parts]), s0)
