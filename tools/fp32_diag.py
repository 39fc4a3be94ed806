"""Developer diagnostic: where do fp32 trajectories leave the fp64 oracle?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import oracle
from pyphy_engine_b200 import sampler
from pyphy_engine_b200.engine import WorldBatch

n, nb, H = 4096, 4, 60
W = sampler.sample_rect_worlds(n, nb, seed=11)
wb = WorldBatch(n, nb, 4, dtype="f32", event_capacity=4_000_000, max_substeps=4096)
wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
traj, _ = wb.step(H, record_traj=True)
traj = traj.cpu().numpy().astype(np.float64)
ev, _ = wb.read_events()
ob = oracle.OracleBatch(W["wall"], W["pos"], W["vel"], W["rad"], W["mass"], trig_mode=1, max_substeps=4096)
otraj, oev = ob.step(H, record_traj=True, max_events=4_000_000)
err_t = np.abs(traj - otraj).max(axis=(2, 3))       # (H, n)
for h in (5, 10, 20, 30, 40, 50, 60):
    e = err_t[:h].max(0)
    print("horizon %2d: p50 %.2e p90 %.2e p99 %.2e p99.9 %.2e max %.2e  frac>0.05: %.4f  frac>0.25: %.4f" % (
        h, *np.percentile(e, [50, 90, 99, 99.9]), e.max(), (e > 0.05).mean(), (e > 0.25).mean()))
bad = np.nonzero(err_t.max(0) > 0.25)[0]
print("bad worlds", len(bad))
nbb_o = np.bincount(oev[oev[:, 3] >= 4][:, 0], minlength=n)
nbb_d = np.bincount(ev[ev[:, 3] >= 4][:, 0], minlength=n)
print("ball-ball events per world: all %.2f  bad %.2f" % (nbb_o.mean(), nbb_o[bad].mean()))
print("worlds whose event lists differ:", sum(1 for w in range(n) if not np.array_equal(ev[ev[:,0]==w][:,1:], oev[oev[:,0]==w][:,1:])))
for w in bad[:6]:
    first = int(np.argmax(err_t[:, w] > 0.05))
    print("world", w, "first step err>0.05:", first, "err there", err_t[first, w])
    print("  device events:", ev[ev[:, 0] == w][:, 1:][:12].tolist())
    print("  oracle events:", oev[oev[:, 0] == w][:, 1:][:12].tolist())
st, _ = wb.get_status()
print("device status counts", np.bincount(st), "oracle", np.bincount(ob.status))
