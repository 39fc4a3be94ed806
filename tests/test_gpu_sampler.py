"""Device-side world sampler (ppb_sample_rect_worlds): the DataSaver(isRect=True) distribution drawn
by a CUDA kernel.  It is distribution-equivalent, not stream-equivalent, to the reference, so the
checks are (1) the hard constraints the reference's rejection sampling guarantees, (2) moments
against the host sampler that follows the same recipe, (3) the cage arithmetic against
primitives.create_cage, (4) determinism / independence of the batch split."""
import numpy as np
import pytest

from pyphy_engine_b200 import sampler
from pyphy_engine_b200.engine import WorldBatch
import pyphy_engine_b200.geometry as gm
import pyphy_engine_b200.primitives as pm

pytestmark = pytest.mark.gpu


def _sample(n, nb, seed, dtype="f64", world_index0=0, **kw):
    wb = WorldBatch(n, nb, 4, dtype=dtype, max_substeps=4096)
    nf = wb.sample_rect_worlds(seed=seed, world_index0=world_index0, **kw)
    wall = wb.get_walls()
    pos, vel = wb.get_balls()
    rad, mass = wb.get_radius_mass()
    return wb, nf, wall, pos, vel, rad, mass


def test_constraints_native_arena():
    n, nb = 20000, 4
    wb, nf, wall, pos, vel, rad, mass = _sample(n, nb, 7)
    assert nf == 0
    # cage corner points are the first vertex of every wall; integers + theta2dir noise
    pts = wall[:, :, 0, :]
    hlen = pts[:, 1, 0] - pts[:, 0, 0] + 30
    vlen = pts[:, 2, 1] - pts[:, 1, 1]
    assert np.all((hlen >= 300) & (hlen < 550)) and np.all((vlen >= 300) & (vlen < 550))
    assert np.all(pts[:, 0] >= 0) and np.all(pts[:, 0, 0] + hlen <= 700) and np.all(pts[:, 0, 1] + vlen + 30 <= 700)
    assert np.all(rad == 25) and np.allclose(mass, sampler.ball_mass(25.0))
    assert np.all(pos == np.round(pos))                               # integer centres (dataio.py:365)
    # signed distance to every cage line > r + wThick + 2 (dataio.py:356)
    lo = pts[:, 0, :][:, None, :] + 25 + 30 + 2
    hi = np.stack([pts[:, 1, 0], pts[:, 2, 1]], -1)[:, None, :] - (25 + 30 + 2)
    assert np.all(pos > lo - 1e-9) and np.all(pos < hi + 1e-9)
    d = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1) + np.eye(nb)[None] * 1e9
    assert np.all(d > 50)                                             # no overlap (dataio.py:371)
    speed = np.linalg.norm(vel, axis=-1)
    m = sampler.ball_mass(25.0)
    assert np.all(speed >= 0.5 * 3e4 / m - 1e-9) and np.all(speed <= 0.5 * 8e4 / m + 1e-9)
    wb.close()


def test_cage_arithmetic_matches_create_cage():
    wb, nf, wall, *_ = _sample(64, 1, 3)
    for w in range(64):
        pts = [gm.Point(*wall[w, k, 0]) for k in range(4)]
        ref = np.array([[[v.x(), v.y()] for v in g.bbox_.vert_] for g in pm.create_cage(pts, wThick=30)])
        assert np.allclose(wall[w], ref, rtol=0, atol=1e-9)
    wb.close()


def test_moments_match_the_reference_recipe(tmp_path):
    """Moments of the device sampler vs. the stream-exact DataSaver façade (which reproduces the
    reference's draws bit for bit, tests/test_facade_cpu.py) over a few thousand worlds."""
    import pyphy_engine_b200.dataio as dio
    n_ref, nb = 2500, 3
    ds = dio.DataSaver(rootPath=str(tmp_path), makeDirs=False, randSeed=123, numBalls=nb, mnBallSz=15, mxBallSz=35,
                       mnForce=3e4, mxForce=8e4, wThick=30, isRect=True, mnWLen=300, mxWLen=550, arenaSz=700,
                       mnSeqLen=10, mxSeqLen=10)
    R = dict(xleft=[], ytop=[], rad=[], speed=[], relx=[], absang=[])
    for sp in ds.sample_sequences(n_ref):
        R["xleft"].append(sp.cage_pts[0].x())
        R["ytop"].append(sp.cage_pts[0].y())
        for b in sp.model.world_.dynamic_.values():
            R["rad"].append(b.get_radius())
            R["speed"].append(b.vel_.mag())
            R["relx"].append(b.pos_.x() - sp.cage_pts[0].x())
            R["absang"].append(abs(np.arctan2(b.vel_.y(), b.vel_.x())))
    wb, nf, wall, pos, vel, rad, mass = _sample(40000, nb, 11, mn_ball=15, mx_ball=35)
    D = dict(xleft=wall[:, 0, 0, 0], ytop=wall[:, 0, 0, 1], rad=rad.ravel(), speed=np.linalg.norm(vel, axis=-1).ravel(),
             relx=(pos[:, :, 0] - wall[:, 0, 0, 0][:, None]).ravel(),
             absang=np.abs(np.arctan2(vel[..., 1], vel[..., 0])).ravel())
    for k in R:
        r, d = np.asarray(R[k], np.float64), np.asarray(D[k], np.float64)
        se = r.std() / np.sqrt(len(r)) + d.std() / np.sqrt(len(d))
        assert abs(r.mean() - d.mean()) < 4.5 * se, (k, r.mean(), d.mean(), se)
        assert abs(r.std() - d.std()) < 0.08 * r.std(), (k, r.std(), d.std())
    assert abs((np.arctan2(vel[..., 1], vel[..., 0]) > 0).mean() - 0.5) < 0.01
    wb.close()


def test_deterministic_and_split_independent():
    kw = dict(arena=64, w_thick=3, mn_wlen=50, mx_wlen=58, mn_ball=3, mx_ball=3, mn_force=4.74, mx_force=12.64)
    wb, nf, wall, pos, vel, rad, mass = _sample(3000, 8, 21, dtype="f32", **kw)
    wb2, nf2, wall2, pos2, vel2, *_ = _sample(3000, 8, 21, dtype="f32", **kw)
    assert nf == nf2 and np.array_equal(pos, pos2) and np.array_equal(vel, vel2) and np.array_equal(wall, wall2)
    # the second half drawn as its own batch (a shard) equals the second half of the whole
    wb3, nf3, wall3, pos3, vel3, *_ = _sample(1500, 8, 21, dtype="f32", world_index0=1500, **kw)
    assert np.array_equal(pos3, pos[1500:]) and np.array_equal(vel3, vel[1500:]) and np.array_equal(wall3, wall[1500:])
    other = _sample(3000, 8, 22, dtype="f32", **kw)
    assert not np.array_equal(other[3], pos)
    for b in (wb, wb2, wb3, other[0]):
        b.close()


def test_sampled_worlds_step_and_render():
    kw = dict(arena=64, w_thick=3, mn_wlen=50, mx_wlen=58, mn_ball=3, mx_ball=3, mn_force=4.74, mx_force=12.64)
    n = 8192
    wb = WorldBatch(n, 8, 4, dtype="f32", canvas=(64, 64), max_substeps=256)
    assert wb.sample_rect_worlds(seed=1, **kw) == 0
    wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * 8, [(0, (1, 0, 0, 1))] * 4, 3, 3)
    traj, frames = wb.step(50, record_traj=True, render_every=50)
    status, _ = wb.get_status()
    assert (status == 0).mean() > 0.995
    pos, _ = wb.get_balls()
    wall = wb.get_walls()
    lo = wall[:, 0, 0, :][:, None, :] + 3 + 3 - 0.01
    hi = np.stack([wall[:, 1, 0, 0], wall[:, 2, 0, 1]], -1)[:, None, :] - 3 - 3 + 0.01
    ok = status == 0
    assert ((pos >= lo) & (pos <= hi)).all(axis=(1, 2))[ok].mean() > 0.999
    fr = frames[0].cpu().numpy()
    assert (fr[..., 1] == 0).any() and (fr == 128).any()          # red walls and gray balls were drawn
    wb.close()


def test_too_small_cage_is_reported_not_hung():
    wb = WorldBatch(256, 8, 4, dtype="f32")
    nf = wb.sample_rect_worlds(seed=2, arena=64, w_thick=3, mn_wlen=20, mx_wlen=24, mn_ball=3, mx_ball=3,
                               mn_force=1, mx_force=2, max_tries=50)
    assert nf == 256
    rad, _ = wb.get_radius_mass()
    assert np.all(rad < 0)
    wb.close()


def test_rhombus_tables_against_the_reference_recipe(tmp_path):
    """isRect=False (dataio.py:243-306) on the device: constraints + moments vs the stream-exact DataSaver."""
    import pyphy_engine_b200.dataio as dio
    nb = 3
    kw = dict(arena=1600, w_thick=30, mn_wlen=500, mx_wlen=800, mn_ball=25, mx_ball=25, mn_force=3e4, mx_force=8e4,
              w_theta=[30, 60])
    wb, nf, wall, pos, vel, rad, mass = _sample(30000, nb, 5, **kw)
    assert nf == 0
    pts = wall[:, :, 0, :]
    side = np.linalg.norm(np.roll(pts, -1, axis=1) - pts, axis=-1)
    assert np.allclose(side, side[:, :1], rtol=0, atol=1e-9)                  # a rhombus: four equal sides
    assert np.all((side[:, 0] >= 500) & (side[:, 0] <= 800))
    ang = np.degrees(np.arctan2(-(pts[:, 1, 1] - pts[:, 0, 1]), pts[:, 1, 0] - pts[:, 0, 0]))
    assert np.all(np.isclose(ang, 30, atol=1e-9) | np.isclose(ang, 60, atol=1e-9))
    assert abs(np.isclose(ang, 30, atol=1e-9).mean() - 0.5) < 0.02
    assert wall.min() >= 0 and wall.max() <= 1600
    # every ball keeps r + wThick + 2 from every cage line (signed distance, dataio.py:356) and off the others
    for q in range(4):
        p0, p1 = pts[:, q], pts[:, (q + 1) % 4]
        a = -(p1[:, 1] - p0[:, 1]); b = p1[:, 0] - p0[:, 0]; c = p0[:, 0] * p1[:, 1] - p0[:, 1] * p1[:, 0]
        d = (a[:, None] * pos[..., 0] + b[:, None] * pos[..., 1] + c[:, None]) / np.sqrt(a * a + b * b)[:, None]
        assert np.all(np.abs(d) > 25 + 30 + 2 - 1e-6)
    dd = np.linalg.norm(pos[:, :, None, :] - pos[:, None, :, :], axis=-1) + np.eye(nb)[None] * 1e9
    assert np.all(dd > 50)
    ds = dio.DataSaver(rootPath=str(tmp_path), makeDirs=False, randSeed=9, numBalls=nb, mnBallSz=25, mxBallSz=25,
                       mnForce=3e4, mxForce=8e4, wThick=30, isRect=False, wTheta=[30, 60], mnWLen=500, mxWLen=800,
                       arenaSz=1600, mnSeqLen=10, mxSeqLen=10)
    R = dict(x0=[], y0=[], side=[], relx=[])
    for sp in ds.sample_sequences(1500):
        R["x0"].append(sp.cage_pts[0].x()); R["y0"].append(sp.cage_pts[0].y())
        R["side"].append((sp.cage_pts[1] - sp.cage_pts[0]).mag())
        for b in sp.model.world_.dynamic_.values():
            R["relx"].append(b.pos_.x() - sp.cage_pts[0].x())
    D = dict(x0=pts[:, 0, 0], y0=pts[:, 0, 1], side=side[:, 0], relx=(pos[:, :, 0] - pts[:, 0, 0][:, None]).ravel())
    for k in R:
        r, d = np.asarray(R[k]), np.asarray(D[k])
        se = r.std() / np.sqrt(len(r)) + d.std() / np.sqrt(len(d))
        assert abs(r.mean() - d.mean()) < 4.5 * se, (k, r.mean(), d.mean(), se)
        assert abs(r.std() - d.std()) < 0.08 * r.std(), (k, r.std(), d.std())
    # and they simulate: 100 fp64 steps without a reference assertion
    wb.step(100)
    assert (wb.get_status()[0] == 0).mean() > 0.99
    wb.close()
