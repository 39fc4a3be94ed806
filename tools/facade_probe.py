"""Latency of the reference-facing single-world API (BASELINE configs[0], [1])."""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import make_golden as mg
import pyphy_engine_b200.geometry as gm, pyphy_engine_b200.primitives as pm, pyphy_engine_b200.dataio as dio
ns = types.SimpleNamespace(gm=gm, pm=pm, dio=dio)
for name, build in (("config1_64 (1 ball, 64x64)", mg.build_config1_64), ("trial_two (2 balls, 640x480)", mg.build_cairo_trial_two)):
    model = build(ns)
    model.step(); model.generate_image()
    t0 = time.perf_counter()
    for _ in range(200): model.step()
    t1 = time.perf_counter()
    for _ in range(200):
        model.step(); im = model.generate_image()
    t2 = time.perf_counter()
    print("%s: step %.1f us, step+generate_image %.1f us" % (name, 1e6 * (t1 - t0) / 200, 1e6 * (t2 - t1) / 200))
import tempfile
kw = dict(mg.RECT, numBalls=4, mnForce=3e4, mxForce=8e4, mnSeqLen=100, mxSeqLen=100)
ds = dio.DataSaver(rootPath=tempfile.mkdtemp(), randSeed=7, makeDirs=False, **kw)
t0 = time.perf_counter(); ims = ds.fetch(); t1 = time.perf_counter()
print("DataSaver.fetch 4 balls, 700x700, 100 frames: %.1f ms (%.0f frames/s)" % (1e3 * (t1 - t0), 100 / (t1 - t0)))
t0 = time.perf_counter(); specs, pos, fr = ds.fetch_batch(32); t1 = time.perf_counter()
print("DataSaver.fetch_batch 32 sequences x 100 frames 700x700: %.1f ms (%.0f frames/s)" % (1e3 * (t1 - t0), 3200 / (t1 - t0)))
