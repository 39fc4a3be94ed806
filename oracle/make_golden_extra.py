"""Extra reference runs that pin the ORACLE only (tests/golden/oracle_extra_golden.npz; CPU tests).

    python oracle/make_golden_extra.py

TEST INFRASTRUCTURE.  Larger ball counts and table shapes than tests/golden/physics_golden.npz holds -- the
configurations BASELINE.json sweeps (16 and 32 balls, polygonal tables) -- stepped by the UNMODIFIED reference
through oracle/ref_shim.py.  The horizons
are moderate; worlds on which the reference itself raises keep their failure outcome (status, step).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import GOLDEN_DIR, build_datasaver, build_ngon, run_case  # noqa: E402


def main():
    ns = ref_shim.load()
    tmpdir = tempfile.mkdtemp(prefix="pyphy_golden_extra_")
    out, cases = {}, []

    def put(name, rec, extra=None):
        for k, v in rec.items():
            out["%s/%s" % (name, k)] = v
        for k, v in (extra or {}).items():
            out["%s/%s" % (name, k)] = v
        cases.append(name)
        print(name, "status", int(rec["status"]), "events", len(rec["events"]), flush=True)

    # BASELINE configs[4]: 16 and 32 balls on rectangular tables (the sweep's arenas)
    for nb, arena, wl, seeds, steps in ((16, 1600, (1000, 1400), tuple(range(900, 906)), 200), (32, 2400, (1600, 2200), tuple(range(910, 914)), 150)):
        for seed in seeds:
            kw = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=arena, mnWLen=wl[0], mxWLen=wl[1],
                      numBalls=nb, mnForce=3e4, mxForce=8e4, mnSeqLen=steps, mxSeqLen=steps)
            model, extra, seq_len = build_datasaver(ns, seed, tmpdir, **kw)
            put("rect%d_s%d" % (nb, seed), run_case(ns, model, seq_len), extra)
    # mixed radii, 6 balls
    for seed in (920, 921, 922):
        kw = dict(wThick=30, isRect=True, mnBallSz=12, mxBallSz=40, arenaSz=1200, mnWLen=700, mxWLen=1000,
                  numBalls=6, mnForce=3e4, mxForce=8e4, mnSeqLen=120, mxSeqLen=120)
        model, extra, seq_len = build_datasaver(ns, seed, tmpdir, **kw)
        put("mixed6_s%d" % seed, run_case(ns, model, seq_len), extra)
    # rhombus tables with 8 balls
    for seed in (930, 931):
        kw = dict(wThick=30, isRect=False, wTheta=[30, 60], mnBallSz=20, mxBallSz=30, arenaSz=2000, mnWLen=800,
                  mxWLen=1100, numBalls=8, mnForce=3e4, mxForce=8e4, mnSeqLen=100, mxSeqLen=100)
        model, extra, seq_len = build_datasaver(ns, seed, tmpdir, **kw)
        put("rhomb8_s%d" % seed, run_case(ns, model, seq_len), extra)
    # regular polygons with more balls, one with gravity switched on
    for n_sides, n_balls, seed in ((6, 8, 940), (7, 5, 941), (12, 6, 942)):
        model = build_ngon(ns, n_sides, n_balls, seed, radius=330.0)
        put("ngon%d_b%d_s%d" % (n_sides, n_balls, seed), run_case(ns, model, 150))
    gm = ns.gm
    model = build_ngon(ns, 6, 4, 943)
    model.set_g(gm.Point(0, 600))
    put("ngon6_gravity_s943", run_case(ns, model, 150))
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "oracle_extra_golden.npz"), **out)
    print("wrote %d cases" % len(cases))


if __name__ == "__main__":
    main()
