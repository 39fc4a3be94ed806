"""Times the UNMODIFIED Python reference's physics path (Dynamics.step, /root/reference/primitives.py:628-656,
driven like dataio.py:118-128) on this container's host cores: one process, and os.cpu_count() processes with one
job per worker the way /root/reference/parallel_fetch.py:12-25 does it.

    python tools/time_reference.py            # writes profiles/r2_python_reference.json

The reference only exists in the build container (Python 2 sources behind oracle/ref_shim.py's import hook; pycairo
is absent, so World.generate_image is a no-op and the figures are physics only): this script cannot run on the GPU
box.  bench.py echoes the committed JSON as cpu_baseline.python_reference so that the reference's own speed stands
next to the C port's, with the machine it was measured on named.
"""
import contextlib
import io
import json
import multiprocessing as mp
import os
import platform
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

NATIVE4 = dict(wThick=30, isRect=True, mnForce=3e4, mxForce=8e4, mnWLen=300, mxWLen=550, mnBallSz=25, mxBallSz=25,
               arenaSz=700, numBalls=4)                                      # save_rect_arena, dataio.py:435-446


def _scaled64(nb, steps):
    from make_golden_64 import scaled64_kwargs
    kw = scaled64_kwargs(nb)
    kw["mnSeqLen"] = kw["mxSeqLen"] = steps
    return kw


def _job(args):
    """One worker: build `n_worlds` worlds with consecutive seeds, step each `steps` times; returns (ball-steps, s)."""
    kind, seed0, n_worlds, steps = args
    import ref_shim
    ns = ref_shim.load()
    from make_golden import build_config1_64, build_datasaver
    tmp = tempfile.mkdtemp(prefix="pyphy_time_")
    models = []
    with contextlib.redirect_stdout(io.StringIO()):
        for w in range(n_worlds):
            if kind == "config1":
                models.append((build_config1_64(ns), 1))
            elif kind == "config2":
                models.append((build_datasaver(ns, seed0 + w, tmp, **_scaled64(4, steps))[0], 4))
            elif kind == "config3":
                models.append((build_datasaver(ns, seed0 + w, tmp, mnSeqLen=steps, mxSeqLen=steps, **NATIVE4)[0], 4))
            elif kind == "config4":
                models.append((build_datasaver(ns, seed0 + w, tmp, **_scaled64(8, steps))[0], 8))
    t0 = time.perf_counter()
    done = 0
    with contextlib.redirect_stdout(io.StringIO()):
        for model, nb in models:
            try:
                for _ in range(steps):
                    model.step()
                    done += nb
            except Exception:        # the reference raises in degenerate contacts (Appendix B Q5/Q12): world ends there
                pass
            ns.events.clear()
    return done, time.perf_counter() - t0


def measure(kind, n_worlds, steps, procs):
    per = max(1, n_worlds // procs)
    jobs = [(kind, 1000 + p * per, per, steps) for p in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_job(jobs[0])]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_job, jobs)
    wall = time.perf_counter() - t0
    ball_steps = sum(r[0] for r in res)
    step_time = max(r[1] for r in res)          # stepping only (world construction excluded), slowest worker
    return dict(config=kind, processes=procs, worlds=per * procs, steps=steps, ball_steps=ball_steps,
                step_seconds=round(step_time, 3), wall_seconds=round(wall, 3),
                ball_steps_per_sec=ball_steps / step_time)


def main():
    ncpu = os.cpu_count()
    plan = [("config1", 1, 100), ("config2", 1, 200), ("config3", 256, 40), ("config4", 64, 40)]
    rows = []
    for kind, n_worlds, steps in plan:
        for procs in (1, ncpu):
            nw = n_worlds if n_worlds > 1 else procs       # configs 1-2: one world per process
            if procs == 1 and n_worlds > 1:
                nw = max(8, n_worlds // ncpu)              # the single-process leg steps 1/ncpu of the sample
            row = measure(kind, nw, steps, procs)
            rows.append(row)
            print(json.dumps(row), flush=True)
    out = dict(what="UNMODIFIED pulkitag/pyphy-engine Dynamics.step through oracle/ref_shim.py, physics only "
                    "(pycairo absent: generate_image is a no-op)",
               host=dict(cpu_count=ncpu, machine=platform.machine(), python=platform.python_version(),
                         cpu_model=next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if "model name" in l), "?"),
                         where="build container (the reference does not travel to the GPU box)"),
               rows=rows)
    with open(os.path.join(ROOT, "profiles", "r2_python_reference.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
