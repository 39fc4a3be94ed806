#!/usr/bin/env python
"""bench.py -- throughput of the simulate-and-render hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the configuration the 1/2/4/8-GPU metric is quoted on):
per GPU 131,072 independent 8-ball worlds on a 64x64 canvas, synthetic random-initialised
rectangular tables; one "step" = Dynamics.step() + World.generate_image() for every world of
the shard (one advance launch per 64 steps + one render launch per step).  Weak scaling: the per-GPU shard is
fixed, 8 GPUs = the full 1,048,576-world config.  Worlds are independent, so there is no
collective in the step loop; NCCL only carries the final timing/statistics reduction.

Prints ONE JSON line on rank 0 (see the field notes in DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORLDS_PER_GPU = 131072
N_BALLS = 8
CANVAS = 64
SEED0 = 20161
MAX_SUBSTEPS = 64
# algorithmic bytes (fp32 state, RGBA8 frames), SURVEY.md 8(d) / DESIGN.md "Kernels"
BYTES_K3_PER_FRAME = CANVAS * CANVAS * 4


def _bytes_advance_per_launch(n_worlds, n_balls, n_steps, n_snapshots):
    """What one advance launch materialises in HBM (a multi-step launch divides by its own bytes):
    per world the 16-byte header read and written, per ball pos + vel + tCol + {radius, mass} read (28 B) and
    pos + tCol written (12 B; velocities, nrmlCol / futureVel and objCol only for worlds that had an event),
    plus 8 B per ball and position snapshot."""
    return n_worlds * (32 + n_balls * 40) + n_snapshots * n_worlds * n_balls * 8


def load_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json,
    written by tools/make_profiles.py from the same workload), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as fh:
            t = json.load(fh)["kernels"]
        key = {"advance": "advance_kernel", "render": "render_tile_kernel"}[kernel]
        return float(t[key]["dram_bytes"])
    except Exception:
        return None


def load_store_ceiling():
    """GB/s a store-only kernel with the render kernel's tile hand-out reaches over the same 2 GiB (measured once on
    B200, profiles/store_ceiling.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "store_ceiling.json")) as fh:
            return float(json.load(fh)["store_only_ceiling_gbs"])
    except Exception:
        return None


def load_python_reference():
    """The UNMODIFIED Python reference timed in the build container (tools/time_reference.py; it cannot travel to the
    GPU box): the committed record, echoed next to the C port's number."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_python_reference.json")) as fh:
            rec = json.load(fh)
        rows = {(r["config"], r["processes"]): r for r in rec["rows"]}
        ncpu = rec["host"]["cpu_count"]
        pick = lambda cfg, procs: round(rows[(cfg, procs)]["ball_steps_per_sec"], 1) if (cfg, procs) in rows else None
        return {"what": rec["what"], "host": rec["host"], "unit": "ball-steps/s",
                "config4_sample_1_process": pick("config4", 1), "config4_sample_%d_processes" % ncpu: pick("config4", ncpu),
                "config3_sample_1_process": pick("config3", 1), "config3_sample_%d_processes" % ncpu: pick("config3", ncpu),
                "config1_1_process": pick("config1", 1), "config2_1_process": pick("config2", 1),
                "source": "profiles/r2_python_reference.json"}
    except Exception as e:           # the record is optional
        return {"unavailable": repr(e)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def make_worlds(first, count, n_balls, seed0=None):
    """Worlds first..first+count-1 of the synthetic job: world w is the same whatever the rank count."""
    from pyphy_engine_b200 import sampler, sharding
    return sharding.sample_world_range(first, count, n_balls, seed0=SEED0 if seed0 is None else seed0,
                                       **sampler.scaled64_params(n_balls))


STYLE_BALL = (1, (0.5, 0.5, 0.5, 1.0), (0.0, 0.0, 0.0, 1.0))   # sThick 1, gray fill, black ring
STYLE_WALL = (0, (1.0, 0.0, 0.0, 1.0))                          # GenericWall, red


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port (C restatement of the reference) on host cores
# --------------------------------------------------------------------------------------
class CpuPort:
    """The C port (oracle/) stepping + rendering a bounded sample of the workload on host cores."""

    def __init__(self, n_worlds, n_balls, threads, seed):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        self.oracle = oracle
        self.W = make_worlds(0, n_worlds, n_balls, seed)
        self.ob = oracle.OracleBatch(self.W["wall"], self.W["pos"], self.W["vel"], self.W["rad"], self.W["mass"],
                                     trig_mode=0, max_substeps=MAX_SUBSTEPS)
        self.threads = threads
        self.n_worlds, self.n_balls = n_worlds, n_balls

    def step(self, render=True):
        """one Dynamics.step() + generate_image() per world; returns seconds."""
        t0 = time.perf_counter()
        self.ob.step(1, n_threads=self.threads)
        if render:
            self.oracle.render_frames(self.W["wall"], self.ob.pos, self.W["rad"], (CANVAS, CANVAS),
                                      s_thick=STYLE_BALL[0], fill=STYLE_BALL[1], stroke=STYLE_BALL[2],
                                      n_threads=self.threads)
        return time.perf_counter() - t0


def cpu_baseline(threads, target_seconds=12.0, n_worlds=4096):
    port = CpuPort(n_worlds, N_BALLS, threads, SEED0)
    for _ in range(3):
        t1 = port.step()
    n_steps = int(min(2000, max(5, target_seconds / max(t1, 1e-5))))
    t = sum(port.step() for _ in range(n_steps))
    live = int((port.ob.status == 0).sum())
    return {"value": live * N_BALLS * n_steps / t, "unit": "ball-steps/s", "cores": threads, "kind": "port",
            "frames_per_sec": live * n_steps / t, "python_reference": load_python_reference(),
            "sample": "%d worlds x %d balls x %d steps after 3 warm-up steps, step+render per step, OpenMP over worlds "
                      "(C restatement of the reference; the Python reference cannot travel to the GPU box)"
                      % (n_worlds, N_BALLS, n_steps)}


def run_reference(args, rank, world_size):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_worlds = 4096
    port = CpuPort(sample_worlds, N_BALLS, threads, SEED0)
    for _ in range(max(args.warmup, 1)):
        port.step()
    t_total = sum(port.step() for _ in range(args.steps))
    live = int((port.ob.status == 0).sum())
    value = live * N_BALLS * args.steps / t_total
    line = {
        "impl": "reference", "metric": "ball_steps_per_sec", "value": value, "unit": "ball-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "frames_per_sec": live * args.steps / t_total,
        "config": {"workload": "configs[3] shard: 131072 worlds/GPU x 8 balls, 64x64 frame rendered every step",
                   "sample": "each timed step = one step+render of a %d-world sample of that workload" % sample_worlds},
        "cpu_baseline": {"value": value, "unit": "ball-steps/s", "cores": threads, "kind": "port",
                         "sample": "%d worlds x %d balls, step+render per timed step, %d steps, OpenMP over worlds; "
                                   "C restatement of the reference (the Python reference cannot travel to the GPU box)"
                                   % (sample_worlds, N_BALLS, args.steps)},
        "e2e": {"value": value, "unit": "ball-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world_size):
    import torch
    import torch.distributed as dist
    from pyphy_engine_b200.engine import WorldBatch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from pyphy_engine_b200 import sharding as _sh
    full_affinity = os.sched_getaffinity(0)
    numa_node = _sh.bind_to_gpu_numa_node(local_rank) if world_size > 1 else None
    distributed = world_size > 1
    if distributed and not dist.is_initialized():
        # NCCL prints its version banner on stdout at the first communicator; keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    # the last rendezvous waits for rank 0's CPU-baseline run: on gloo the other ranks sleep instead of spinning
    cpu_group = dist.new_group(backend="gloo") if distributed else None

    def barrier():
        if distributed:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    from pyphy_engine_b200 import sharding
    first, n_worlds = sharding.shard_range(WORLDS_PER_GPU * world_size, rank, world_size)
    W = make_worlds(first, n_worlds, N_BALLS)                  # shard `rank` of the global world list
    wb = WorldBatch(n_worlds, N_BALLS, 4, dtype="f32", canvas=(CANVAS, CANVAS), device=local_rank,
                    max_substeps=MAX_SUBSTEPS)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.set_styles([STYLE_BALL] * N_BALLS, [STYLE_WALL] * 4, 3, 3)
    # frame ring: 2 slots of 2 GiB; the render of step k (own stream) overlaps the physics of step k+1
    frames = torch.empty((2, n_worlds, CANVAS, CANVAS, 4), dtype=torch.uint8, device=dev)

    # ---- device-resident throughput: W warm-up steps, K timed steps ----
    wb.step_ring_async(args.warmup, frames, 1)
    barrier()
    c0 = wb.counters()
    st0, _ = wb.get_status()
    wb.set_profiling(2)     # events around the render launches only (see the second pass below)
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    wb.step_ring_async(args.steps, frames, 1)      # K steps: K render launches + ceil(K / 64) advance launches
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    # per-kernel CUDA-event times over the timed region (events recorded on the launch stream)
    kms, kn = wb.kernel_times()
    wb.set_profiling(False)
    c1 = wb.counters()
    st1, steps_done = wb.get_status()
    live = int((st1 == 0).sum())
    launches = c1["launches"] - c0["launches"]
    # second, untimed pass of the same K steps with events around the advance launches: an event pair between two
    # kernels of a stream keeps the second from being scheduled while the first drains, so inside the timed region
    # only the dominant kernel is bracketed
    wb.set_profiling(3)
    wb.step_ring_async(args.steps, frames, 1)
    torch.cuda.synchronize(dev)
    kms2, kn2 = wb.kernel_times()
    wb.set_profiling(False)
    kms = [kms2[0], 0.0, kms[2]]
    kn = [kn2[0], 0, kn[2]]

    k_ms, k_n = np.array(kms), np.array(kn)
    avg_us = 1e3 * k_ms / np.maximum(k_n, 1)

    # the render kernel alone (nothing else on the GPU), for comparison with its overlapped time above
    iso0, iso1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wb.render(out=frames[0])
    torch.cuda.synchronize(dev)
    iso0.record()
    for _ in range(5):
        wb.render(out=frames[0])
    iso1.record()
    torch.cuda.synchronize(dev)
    render_alone_us = 1e3 * iso0.elapsed_time(iso1) / 5.0

    # ---- end-to-end through the host-buffer API: H2D of the batch state + step + D2H of positions and frames ----
    # Two sets of pinned host buffers: the call of step i only enqueues (ppb_simulate_host_async); while its frames
    # cross PCIe the inputs of step i+1 are uploaded (the other direction) and its kernels run; step i's buffers are
    # read on the host once ppb_host_sync(keep_in_flight=1) says they are complete.
    traj_hosts = [torch.empty((1, n_worlds, N_BALLS, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    frames_hosts = [torch.empty((1, n_worlds, CANVAS, CANVAS, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    traj_host, frames_host = traj_hosts[0], frames_hosts[0]
    e2e_steps = max(3, min(args.steps, 10))
    h2d = W["pos"].nbytes // 2 + W["vel"].nbytes // 2 + W["rad"].nbytes + W["wall"].nbytes // 2   # fp32 on the wire
    d2h = traj_host.numel() * 4 + frames_host.numel()
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.simulate_host(1, traj_hosts[0], frames_hosts[0], 1)        # warm-up (allocates the staging buffers)
    wb.simulate_host(1, traj_hosts[1], frames_hosts[1], 1)
    barrier()
    e2e_checksum = 0
    cp0 = wb.counters()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        wb.set_walls(W["wall"])                                       # host -> device: this step's input tables
        wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        wb.simulate_host(1, traj_hosts[i & 1], frames_hosts[i & 1], 1, wait=False)   # step + render + device -> host, enqueued
        wb.host_sync(keep_in_flight=1)                                # step i-1 is on the host now
        if i > 0:
            e2e_checksum += int(frames_hosts[(i - 1) & 1][0, 0, 0, :8].sum()) + int(traj_hosts[(i - 1) & 1][0, 0, 0, 0] != 0)
    wb.host_sync()
    e2e_checksum += int(frames_hosts[(e2e_steps - 1) & 1][0, :64].sum())
    barrier()
    e2e_s = time.perf_counter() - t0
    cp1 = wb.counters()
    traj_host = traj_hosts[(e2e_steps - 1) & 1]
    # bytes that crossed PCIe for the frames (run-length tokens + per-frame table; raw when a batch does not compress)
    d2h_wire = (cp1["d2h_packed_bytes"] - cp0["d2h_packed_bytes"] + cp1["d2h_raw_bytes"] - cp0["d2h_raw_bytes"]) / e2e_steps \
        + traj_host.numel() * 4
    unpack_ms = 1e-3 * (cp1["unpack_us"] - cp0["unpack_us"]) / e2e_steps

    # ---- reductions: max time over ranks, totals over ranks ----
    stats = torch.tensor([ms, e2e_s, float(live), float(c1["events"] - c0["events"]), float(launches)],
                         dtype=torch.float64, device=dev)
    if distributed:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[1])
        live_total, events_total, launches_total = float(sm[2]), float(sm[3]), float(sm[4])
    else:
        live_total, events_total, launches_total = float(live), float(stats[3]), float(launches)

    # ---- end-of-run gather of the datagen shard (NCCL over NVLink when distributed): 100 steps of positions of every
    #      rank's worlds arrive on rank 0 in world order (`position` (2N, seqLen) fp32 per world).  The
    #      first gather of a process pays the lazy NCCL connection set-up and the allocator's first cudaMallocs, so an
    #      untimed gather of the same shard goes first; not part of any timed region above ----
    GATHER_STEPS = 100
    shard = torch.empty((GATHER_STEPS, n_worlds, N_BALLS, 2), dtype=torch.float32, device=dev)
    wb.step_async(GATHER_STEPS, traj=shard)
    torch.cuda.synchronize(dev)
    warm = _sh.gather_trajectories(shard, n_worlds * world_size, dst=0)      # warm-up: connections, allocator
    del warm
    torch.cuda.synchronize(dev)
    barrier()
    gt = {}
    tg0 = time.perf_counter()
    gathered = _sh.gather_trajectories(shard, n_worlds * world_size, dst=0, timing=gt)
    torch.cuda.synchronize(dev)
    barrier()
    gather_ms = 1e3 * (time.perf_counter() - tg0)
    gather_info = None
    if rank == 0:
        nbytes = int(gathered.numel() * 4)
        coll_ms = gt.get("collective_ms")
        gather_info = {"shape": list(gathered.shape), "bytes": nbytes, "ms": gather_ms, "collective_ms": coll_ms,
                       "gb_per_s": nbytes * (world_size - 1) / max(world_size, 1) / (coll_ms * 1e-3) / 1e9
                       if distributed and coll_ms else None,
                       "backend": "nccl (dist.gather to rank 0; collective_ms = CUDA events around the collective on rank 0, "
                                  "GB/s = bytes arriving from the other ranks / collective_ms; ms = wall clock of the whole "
                                  "call between barriers, incl. the reordering into world order on rank 0)"
                                  if distributed else "none (single rank)",
                       "checksum": float(gathered[-1].double().sum().item())}
    del gathered, shard

    # ---- the same workload in the bit-exact fp64 mode and under the default sub-step guard (own sub-records; the
    #      headline `value` is the fp32 run above with the bench's guard of 64) ----
    def variant(dtype, max_substeps, k_steps, w_steps):
        vb = WorldBatch(n_worlds, N_BALLS, 4, dtype=dtype, canvas=(CANVAS, CANVAS), device=local_rank,
                        max_substeps=max_substeps)
        vb.set_walls(W["wall"])
        vb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        vb.set_styles([STYLE_BALL] * N_BALLS, [STYLE_WALL] * 4, 3, 3)
        vb.step_ring_async(w_steps, frames, 1)
        torch.cuda.synchronize(dev)
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        vb.step_ring_async(k_steps, frames, 1)
        v1.record()
        torch.cuda.synchronize(dev)
        vms = v0.elapsed_time(v1)
        vst, _ = vb.get_status()
        vlive = int((vst == 0).sum())
        vb.close()
        return {"dtype": dtype, "max_substeps": max_substeps, "steps": k_steps, "warmup": w_steps,
                "ms_per_step": vms / k_steps, "ball_steps_per_sec_rank0": vlive * N_BALLS * k_steps / (vms * 1e-3),
                "frames_per_sec_rank0": vlive * k_steps / (vms * 1e-3), "halted_worlds_rank0": n_worlds - vlive}
    f64_mode = variant("f64", MAX_SUBSTEPS, 5, 3) if rank == 0 else None
    default_guard = variant("f32", 1 << 20, args.steps, args.warmup) if rank == 0 else None

    # ---- BASELINE.json configs[2]: 65,536 independent 4-ball worlds, random init, physics only, one GPU (rank 0;
    #      the reference's own table: DataSaver / save_rect_arena distribution, arena 700, r = 25, dataio.py:435-446) ----
    def physics_only(nw, nb, k_steps, w_steps):
        pb = WorldBatch(nw, nb, 4, dtype="f32", device=local_rank, max_substeps=MAX_SUBSTEPS)
        nf = pb.sample_rect_worlds(seed=3, arena=700, mn_wlen=300, mx_wlen=550)
        pb.step_async(w_steps)
        torch.cuda.synchronize(dev)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pb.step_async(k_steps)
        p1.record()
        torch.cuda.synchronize(dev)
        pms = p0.elapsed_time(p1)
        pst, _ = pb.get_status()
        plive = int((pst == 0).sum())
        pb.close()
        return {"workload": "configs[2]: %d worlds x %d balls, physics only, fp32, DataSaver rect tables (arena 700, r = 25)" % (nw, nb),
                "steps": k_steps, "warmup": w_steps, "ms_per_step": pms / k_steps, "sampler_failures": int(nf),
                "ball_steps_per_sec_rank0": plive * nb * k_steps / (pms * 1e-3), "halted_worlds_rank0": nw - plive,
                "max_substeps": MAX_SUBSTEPS}
    physics_cfg2 = physics_only(65536, 4, 1000, 100) if rank == 0 else None

    if rank == 0:
        total_worlds = n_worlds * world_size
        ball_steps = live_total * N_BALLS * args.steps          # halted worlds do not count
        value = ball_steps / (ms * 1e-3)
        fps = live_total * args.steps / (ms * 1e-3)
        peak, peak_src = load_peaks()
        # dominant kernel of this workload: the one that moves the bytes and the time (render: 99 % of the step's
        # algorithmic traffic, about 90 % of its duration)
        names = {0: "advance", 2: "render"}
        steps_per_adv = args.steps / max(kn[0], 1)
        alg_bytes = {0: _bytes_advance_per_launch(n_worlds, N_BALLS, steps_per_adv, steps_per_adv),
                     2: BYTES_K3_PER_FRAME * n_worlds}
        achieved = {i: alg_bytes[i] / (avg_us[i] * 1e-6) / 1e9 if avg_us[i] > 0 else 0.0 for i in names}
        dom = 2
        os.sched_setaffinity(0, full_affinity)     # the CPU baseline may use every host core again
        cpu = cpu_baseline(len(full_affinity))
        line = {
            "metric": "ball_steps_per_sec", "value": value, "unit": "ball-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_sec": fps,
            "config": {"workload": "configs[3] shard: 131072 worlds/GPU x 8 balls, 64x64 frame rendered every step",
                       "worlds_total": total_worlds, "balls_per_world": N_BALLS, "canvas": [CANVAS, CANVAS],
                       "l2": "per-step working set (2 GiB of frames + 70 MB of state per GPU) exceeds the 126 MB L2",
                       "schedule": "one advance launch per chunk of up to 64 steps leaves a position snapshot per step, the chunk's render launches follow on their own stream; the next advance launch starts beside the last 2 of them; frames into a 2-slot ring",
                       "max_substeps": MAX_SUBSTEPS, "halted_worlds": int(total_worlds - live_total),
                       "events_per_world_step": events_total / (total_worlds * max(args.steps, 1)),
                       "pair_tests_per_sec_rank0": (c1["pair_tests"] - c0["pair_tests"]) / (ms * 1e-3)},
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved[dom], "peak": peak,
                         "unit": "GB/s", "frac": achieved[dom] / peak, "traffic": load_traffic(names[dom]), "peak_source": peak_src,
                         "avg_launch_us": float(avg_us[dom]),
                         "note": "render: CUDA-event times of every launch inside the timed region (the last 2 of a chunk share the SMs "
                                 "with the next chunk's advance launch); advance: event times of a second pass of the same K steps, "
                                 "%.1f steps per launch, bytes = what that launch materialises (state round trip + snapshots) -- "
                                 "it is bound by instruction issue and by its longest sub-step chain, not by HBM; "
                                 "render alone: %.1f us/launch = %.3f of peak"
                                 % (steps_per_adv, render_alone_us, alg_bytes[2] / (render_alone_us * 1e-6) / 1e9 / peak),
                         "render_alone_us": render_alone_us,
                         # the peak above is a copy (read + write) figure; a write-only stream goes above it, so the
                         # fraction can exceed 1: the store-only ceiling of this launch shape is reported beside it
                         "store_only_ceiling": (lambda c: None if not c else {
                             "gbs": c, "frac": achieved[dom] / c,
                             "source": "profiles/store_ceiling.json (tools/store_pattern.cu: the render kernel's bulk-store "
                                       "skeleton and tile hand-out without any drawing or state loads)"})(load_store_ceiling()),
                         "all_kernels": {names[i]: {"avg_launch_us": float(avg_us[i]), "achieved_gbs": achieved[i],
                                                    "frac": achieved[i] / peak, "algorithmic_bytes_per_launch": alg_bytes[i]}
                                         for i in names}},
            "cpu_baseline": cpu,
            "e2e": {"value": live_total * N_BALLS * e2e_steps / e2e_s, "unit": "ball-steps/s",
                    "frames_per_sec": live_total * e2e_steps / e2e_s,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_wire),
                    "delivered_bytes_per_step": int(d2h), "unpack_ms_per_step_rank0": unpack_ms, "steps": e2e_steps,
                    "api": "per step: WorldBatch.set_walls/set_balls (H2D) + simulate_host(wait=False) (ppb_simulate_host_async: step, render, "
                           "run-length pack on the device, D2H of the tokens, RGBA rebuilt in the caller's pinned array by the library's host "
                           "threads) + host_sync(keep_in_flight=1); two sets of pinned host buffers, so the upload of step i+1 overlaps the "
                           "download of step i; d2h_bytes_per_step = bytes that crossed PCIe, delivered_bytes_per_step = bytes of "
                           "positions + RGBA frames the caller receives",
                    "numa_node_rank0": numa_node},
            "final_gather": gather_info,
            "f64_mode": f64_mode,
            "default_guard": default_guard,
            "physics_only_configs2": physics_cfg2,
            "gpu_launches": int(launches_total),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        }
        print(json.dumps(line), flush=True)
    if distributed:
        torch.cuda.synchronize(dev)
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world_size)
    else:
        run_ours(args, rank, local_rank, world_size)


if __name__ == "__main__":
    main()
