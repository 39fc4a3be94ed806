"""TEST INFRASTRUCTURE ONLY -- reference runs of the 64-px tables the benchmark steps.

    python oracle/make_golden_64.py        # rewrites tests/golden/physics64_golden.npz

SURVEY.md 8(d) config 2 (4 balls) and config 4 (8 balls) at 64 px are the native ICLR16 rectangular arena
(/root/reference/dataio.py:435-446) with every length scaled by 64/700: arena 64, wThick 3, radius 3, sThick 1,
cage sides 36-50 px (4 balls) / 50-58 px (8 balls), forces 4.74-12.64 so that speeds scale alike
(`pyphy_engine_b200.sampler.scaled64_params`, the parameters bench.py samples its worlds from).  Here the
UNMODIFIED reference draws and steps such worlds itself -- `DataSaver(...).generate_model()` then 200 x
`Dynamics.step()` through oracle/ref_shim.py -- so that the absolute constants the reference carries (the
[R - 0.1, R] overlap nudge and its 1e-6, geometry.py:411-413; TOL = 1e-8, geometry.py:7), all tuned at R = 50,
are pinned at R = 6 as well.  Also one world per ball count with 12 balls' worth of crowding is not needed: the
8-ball cage is already at the benchmark's density.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import GOLDEN_DIR, build_datasaver, run_case  # noqa: E402

N_STEPS = 200
SEEDS = {4: tuple(range(6400, 6420)), 8: tuple(range(6800, 6820))}


def scaled64_kwargs(n_balls):
    """DataSaver keyword arguments of the 64-px arena == sampler.scaled64_params(n_balls)."""
    wl = (36, 50) if n_balls <= 4 else (50, 58)
    return dict(wThick=3, isRect=True, mnBallSz=3, mxBallSz=3, arenaSz=64, mnWLen=wl[0], mxWLen=wl[1],
                numBalls=n_balls, mnForce=4.74, mxForce=12.64, mnSeqLen=N_STEPS, mxSeqLen=N_STEPS)


def main():
    ns = ref_shim.load()
    tmpdir = tempfile.mkdtemp(prefix="pyphy_golden64_")
    out, cases = {}, []
    for nb, seeds in SEEDS.items():
        for seed in seeds:
            model, extra, seq_len = build_datasaver(ns, seed, tmpdir, **scaled64_kwargs(nb))
            rec = run_case(ns, model, seq_len)
            name = "rect64_b%d_s%d" % (nb, seed)
            for k, v in rec.items():
                out["%s/%s" % (name, k)] = v
            for k, v in extra.items():
                out["%s/%s" % (name, k)] = v
            cases.append(name)
            print(name, "status", int(rec["status"]), "step", int(rec["status_step"]), "events", len(rec["events"]),
                  flush=True)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "physics64_golden.npz"), **out)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
