"""Host-side multi-GPU logic on CPU: world partitioning, G-independent world sampling and the
end-of-run statistics / trajectory gathers over ``torch.distributed`` (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyphy_engine_b200 import sharding


def test_shard_ranges_partition_the_worlds():
    for n in (0, 1, 7, 4096, 131072, 1048576, 1000003):
        for g in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, g) for r in range(g)]
            assert spans[0][0] == 0
            for a, b in zip(spans, spans[1:]):
                assert a[0] + a[1] == b[0]
            assert spans[-1][0] + spans[-1][1] == n
            assert max(c for _, c in spans) - min(c for _, c in spans) <= -(-n // g)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_world_sampling_is_independent_of_the_rank_count():
    kw = dict(arena=64, w_thick=3, mn_wlen=50, mx_wlen=58, mn_ball=3, mx_ball=3, mn_force=4.74, mx_force=12.64)
    n = 3 * sharding.BLOCK // 2
    full = sharding.sample_world_range(0, n, 4, seed0=11, **kw)
    for g in (2, 3):
        parts = [sharding.sample_world_range(*sharding.shard_range(n, r, g), 4, seed0=11, **kw) for r in range(g)]
        for k in ("wall", "pos", "vel", "rad", "mass"):
            assert np.array_equal(np.concatenate([p[k] for p in parts], 0), full[k]), k
    assert sharding.sample_world_range(5, 0, 4, seed0=11, **kw)["pos"].shape == (0, 4, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, n_worlds, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        first, count = sharding.shard_range(n_worlds, rank, world_size)
        kw = dict(arena=64, w_thick=3, mn_wlen=50, mx_wlen=58, mn_ball=3, mx_ball=3, mn_force=4.74, mx_force=12.64)
        W = sharding.sample_world_range(first, count, 2, seed0=3, **kw)
        # stand-in for the device step loop: a deterministic per-world function of the initial state
        steps = 3
        traj = np.stack([W["pos"] + (s + 1) * 0.01 * W["vel"] for s in range(steps)]).astype(np.float32)
        stats = sharding.all_gather_stats({"worlds": count, "ball_steps": count * 2 * steps, "ms": 10.0 + rank,
                                           "checksum": float(sharding.world_checksum(W["pos"]).sum() % (1 << 40))})
        full = sharding.gather_trajectories(torch.from_numpy(traj), n_worlds, dst=0)
        if rank == 0:
            np.savez(os.path.join(out_dir, "r0.npz"), traj=full.numpy(), **{k: v for k, v in stats.items()})
        else:
            assert full is None
            assert stats["worlds"].sum() == n_worlds
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 3])
def test_gather_over_gloo(world_size, tmp_path):
    n_worlds = 1000
    port = _free_port()
    mp.spawn(_worker, args=(world_size, port, n_worlds, str(tmp_path)), nprocs=world_size, join=True)
    got = np.load(os.path.join(str(tmp_path), "r0.npz"))
    assert got["worlds"].sum() == n_worlds and len(got["worlds"]) == world_size
    assert got["ball_steps"].sum() == n_worlds * 2 * 3
    assert got["ms"].max() == 10.0 + world_size - 1            # the job time is the max over ranks
    kw = dict(arena=64, w_thick=3, mn_wlen=50, mx_wlen=58, mn_ball=3, mx_ball=3, mn_force=4.74, mx_force=12.64)
    W = sharding.sample_world_range(0, n_worlds, 2, seed0=3, **kw)
    ref = np.stack([W["pos"] + (s + 1) * 0.01 * W["vel"] for s in range(3)]).astype(np.float32)
    assert np.array_equal(got["traj"], ref)                    # shards arrive in world order, nothing lost


def test_single_process_fallbacks():
    assert not sharding.is_distributed()
    s = sharding.all_gather_stats({"a": 1.0, "b": 2.0})
    assert s["a"].tolist() == [1.0] and s["b"].tolist() == [2.0]
    t = torch.zeros(2, 5, 3, 2)
    assert sharding.gather_trajectories(t, 5) is t


def test_parse_topo_table():
    """The nvidia-smi fallback of bind_to_gpu_numa_node reads the affinity columns of the right row."""
    from pyphy_engine_b200.sharding import _parse_topo
    out = ("\x1b[4m\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
           "GPU0\t X \tNV18\tPIX\t0-23,96-119\t0\t\tN/A\n"
           "GPU1\tNV18\t X \tNODE\t24-47\t1\t\tN/A\n"
           "NIC0\tPIX\tNODE\t X \t\t\t\t\n\nLegend:\n\n  X    = Self\n")
    cpus, node = _parse_topo(out, 0)
    assert node == 0 and cpus == set(range(0, 24)) | set(range(96, 120))
    cpus, node = _parse_topo(out, 1)
    assert node == 1 and cpus == set(range(24, 48))
    assert _parse_topo(out, 5) is None
    assert _parse_topo("", 0) is None
