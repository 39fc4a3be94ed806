"""Partitioning of independent worlds over the GPUs of one node.

Worlds never interact (the reference's only parallelism is one ``DataSaver`` per process,
parallel_fetch.py:12-25), so the step loop needs no collective: every rank owns a contiguous
block of world indices and runs the same three kernels on it.  ``torch.distributed`` (NCCL on
the GPUs, gloo in the CPU tests) only carries the end-of-run statistics and, on request, the
recorded trajectories.

World ``w`` of a synthetic job is defined by its index alone (block ``w // BLOCK`` of the sampler
stream ``seed0 + block``), so the same world is produced whatever the number of ranks -- which
is what makes "1 GPU vs N GPUs give identical per-world results" checkable.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

BLOCK = 4096   # worlds per sampler block


def shard_range(n_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first, count) of the worlds of ``rank``: contiguous blocks of ceil(n / world_size)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    per = -(-int(n_worlds) // world_size)
    first = min(int(n_worlds), rank * per)
    return first, max(0, min(int(n_worlds), first + per) - first)


def sample_world_range(first: int, count: int, n_balls: int, seed0: int = 0, **sampler_kw) -> Dict[str, np.ndarray]:
    """Worlds ``first .. first+count-1`` of the synthetic job ``seed0`` (see module docstring)."""
    from . import sampler
    parts = []
    b0, b1 = first // BLOCK, (first + count - 1) // BLOCK if count > 0 else first // BLOCK - 1
    for blk in range(b0, b1 + 1):
        W = sampler.sample_rect_worlds(BLOCK, n_balls, seed=seed0 + blk, **sampler_kw)
        lo = max(first, blk * BLOCK) - blk * BLOCK
        hi = min(first + count, (blk + 1) * BLOCK) - blk * BLOCK
        parts.append({k: v[lo:hi] for k, v in W.items()})
    if not parts:
        W = sampler.sample_rect_worlds(1, n_balls, seed=seed0, **sampler_kw)
        return {k: v[:0] for k, v in W.items()}
    return {k: np.concatenate([p[k] for p in parts], 0) for k in parts[0]}


def bind_to_gpu_numa_node(device: int) -> Optional[int]:
    """Pin this process (and therefore its page-locked staging buffers, allocated first-touch) to the
    CPU cores of the NUMA node the GPU hangs off, so that the frame D2H stream of each rank stays on
    its own socket's memory controllers.  Returns the node, or None when the topology is not exposed."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        if node < 0:
            raise OSError("no NUMA node in sysfs")      # virtual machines report -1: ask nvidia-smi instead
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        pass
    return _bind_from_nvidia_smi(device)


def _parse_cpu_list(spec):
    cpus = set()
    for part in spec.split(","):
        part = part.strip()
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part.isdigit():
            cpus.add(int(part))
    return cpus


def _parse_topo(out: str, device: int):
    """``nvidia-smi topo -m`` text -> (cpu set, numa node or -1) of ``GPU<device>``, or None."""
    import re
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    lines = [l for l in out.splitlines() if l.strip()]
    if not lines:
        return None
    header = re.split(r"\t+|\s{2,}", lines[0].strip())
    if "CPU Affinity" not in header:
        return None
    ci = header.index("CPU Affinity")
    ni = header.index("NUMA Affinity") if "NUMA Affinity" in header else None
    for l in lines[1:]:
        cols = re.split(r"\t+|\s{2,}", l.strip())
        # a row has one more leading column (its own label) than the header
        if cols and cols[0] == "GPU%d" % device and len(cols) > ci + 1:
            node = -1
            if ni is not None and len(cols) > ni + 1 and cols[ni + 1].strip().isdigit():
                node = int(cols[ni + 1])
            return _parse_cpu_list(cols[ci + 1]), node
    return None


def _bind_from_nvidia_smi(device: int) -> Optional[int]:
    """Fallback when sysfs hides the NUMA node (containers): the "CPU Affinity" / "NUMA Affinity" columns of
    ``nvidia-smi topo -m``."""
    import os
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hit = _parse_topo(out, device)
        if hit is None:
            return None
        cpus = hit[0] & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return hit[1]
    except Exception:
        return None


def _dist():
    import torch.distributed as dist
    return dist


def is_distributed() -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_gather_stats(stats: Dict[str, float], device=None) -> Dict[str, np.ndarray]:
    """Every rank contributes a dict of scalars; every rank receives ``{key: array[world_size]}``.
    One small all_gather (a few hundred bytes) -- the only collective of a run."""
    import torch
    keys = sorted(stats)
    local = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    if not is_distributed():
        return {k: np.array([float(stats[k])]) for k in keys}
    dist = _dist()
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local)
    mat = torch.stack(out).cpu().numpy()
    return {k: mat[:, i] for i, k in enumerate(keys)}


def gather_trajectories(local, n_worlds_total: int, dst: int = 0, timing: dict | None = None):
    """Gather the per-rank ``position`` shards (steps, local_worlds, balls, 2) on rank ``dst`` in
    world order (the reference collects ``Pool`` results the same way, parallel_fetch.py:22-25).
    Returns the full tensor on ``dst`` and None elsewhere.  With ``timing`` (a dict) the collective alone is
    bracketed by CUDA events on the current stream and its duration lands in ``timing["collective_ms"]``
    (the padding of ragged shards and the reordering into world order on ``dst`` are outside that bracket)."""
    import torch
    if not is_distributed():
        return local
    dist = _dist()
    rank, ws = dist.get_rank(), dist.get_world_size()
    per = -(-int(n_worlds_total) // ws)
    if local.shape[1] == per and local.is_contiguous():
        pad = local                     # even shards (the usual case): nothing to pad
    else:
        pad = torch.zeros((local.shape[0], per) + tuple(local.shape[2:]), dtype=local.dtype, device=local.device)
        pad[:, :local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)] if rank == dst else None
    ev = None
    if timing is not None and local.is_cuda:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    dist.gather(pad, bufs, dst=dst)
    if ev is not None:
        ev[1].record()
        ev[1].synchronize()
        timing["collective_ms"] = ev[0].elapsed_time(ev[1])
    if rank != dst:
        return None
    parts = []
    for r in range(ws):
        _, cnt = shard_range(n_worlds_total, r, ws)
        parts.append(bufs[r][:, :cnt])
    return torch.cat(parts, 1)


def world_checksum(pos: np.ndarray) -> np.ndarray:
    """Order-independent per-world fingerprint of a position array (..., balls, 2) -> (...,) uint64,
    used to compare shards across different rank counts."""
    b = np.ascontiguousarray(pos, np.float64).view(np.uint64)
    b = b.reshape(b.shape[:-2] + (-1,))
    mult = (np.arange(b.shape[-1], dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1))
    with np.errstate(over="ignore"):
        return (b * mult).sum(axis=-1, dtype=np.uint64)
