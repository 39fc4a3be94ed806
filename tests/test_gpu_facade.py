"""GPU parity tests through the REFERENCE-FACING API (geometry / primitives / dynamics / dataio):
the same builder code that produced the golden fixtures from the unmodified reference
(oracle/make_golden.py) is run against the façade modules, stepped with ``Dynamics.step()`` on the
device, and compared with what the reference did -- collision events identical, positions within
1e-6 px (fp64 batches), failures raised in the same step."""
import os
import types

import numpy as np
import pytest

import make_golden as mg
import oracle
from conftest import GOLDEN_DIR, load_golden
from test_facade_cpu import DS_CASES, HAND_BUILT, datasaver_kwargs

import pyphy_engine_b200.dataio as dio
import pyphy_engine_b200.dynamics as dy
import pyphy_engine_b200.geometry as gm
import pyphy_engine_b200.primitives as pm

pytestmark = pytest.mark.gpu

NS = types.SimpleNamespace(gm=gm, pm=pm, dio=dio)
CASES = {c.name: c for c in load_golden("physics_golden.npz")}
FAIL = {c.name: c for c in load_golden("failure_golden.npz")}
TOC = np.load(os.path.join(GOLDEN_DIR, "toc_golden.npz"))


def run_model(model, n_steps):
    names = model.get_dynamic_object_names()
    traj = np.zeros((n_steps, len(names), 2))
    for s in range(n_steps):
        model.step()
        for j, n in enumerate(names):
            p = model.get_object_position(n)
            traj[s, j] = (p.x(), p.y())
    return traj


def events_as_array(model):
    world = model.world_
    walls = [n for n in world.get_object_names() if n.startswith("wall")]
    balls = [n for n in world.get_object_names() if n.startswith("ball")]
    idx = {n: i for i, n in enumerate(walls)}
    idx.update({n: len(walls) + i for i, n in enumerate(balls)})
    bidx = {n: i for i, n in enumerate(balls)}
    return np.array([(s, bidx[b], idx[p]) for s, b, p in model.get_events()], np.int32).reshape(-1, 3)


@pytest.mark.parametrize("name,builder", HAND_BUILT, ids=[n for n, _ in HAND_BUILT])
def test_hand_built_worlds_step_like_the_reference(name, builder):
    ref = CASES[name]
    model = builder(NS)
    traj = run_model(model, ref.n_steps)
    assert np.array_equal(events_as_array(model), ref["events"])
    assert np.max(np.abs(traj - ref["traj"])) < 1e-6
    vel = np.array([[model.get_object(n).get_velocity().x(), model.get_object(n).get_velocity().y()]
                    for n in model.get_dynamic_object_names()])
    assert np.max(np.abs(vel - ref["vel_final"])) < 1e-6 * max(1.0, np.abs(ref["vel_final"]).max())


@pytest.mark.parametrize("name", [n for n in DS_CASES if not n.startswith("crowd12")][:16] +
                         [n for n in DS_CASES if n.startswith(("opp2", "rhombval"))])
def test_datasaver_worlds_step_like_the_reference(name, tmp_path):
    kw, seed = datasaver_kwargs(name)
    model, extra, seq_len = mg.build_datasaver(NS, seed, str(tmp_path), **kw)
    ref = CASES[name]
    traj = run_model(model, seq_len)
    assert np.array_equal(events_as_array(model), ref["events"])
    assert np.max(np.abs(traj - ref["traj"])) < 1e-6


@pytest.mark.parametrize("kind", ["overlap", "overlap_late", "centre_on_line", "amiss", "corner_clip"])
def test_failures_raise_in_the_same_step(kind):
    ref = FAIL["fail_" + kind]
    model = mg.build_failure(NS, kind)
    done = 3 if kind == "overlap_late" else 0     # the builder already took 3 steps
    with pytest.raises(AssertionError):
        for s in range(done, ref.n_steps + done):
            model.step()
    assert model._nsteps - 1 == ref.status_step + done


def test_toc_ball_wall_against_reference_vectors():
    wi, wo = TOC["wall_in"], TOC["wall_out"]
    n_hit = 0
    for x, ref in list(zip(wi, wo))[:150]:
        wall = pm.GenericWall(gm.Point(x[0], x[1]), gm.Point(x[2], x[3]), wThick=x[4])
        ball = pm.Ball(radius=x[9], initPos=gm.Point(x[5], x[6]), initVel=gm.Point(x[7], x[8]))
        t, nr, pt = dy.get_toc_ball_wall(ball, wall)
        if not np.isfinite(ref[0]):
            assert nr is None and pt is None and not np.isfinite(t)
            continue
        n_hit += 1
        assert t == ref[0]                                   # same fp64 operation order
        assert (nr.x(), nr.y()) == (ref[1], ref[2])
        assert abs(pt.x() - ref[3]) < 1e-9 and abs(pt.y() - ref[4]) < 1e-9
    assert n_hit > 20


def test_toc_ball_ball_against_reference_vectors():
    bi, bo = TOC["ball_in"], TOC["ball_out"]
    n_hit = 0
    for x, ref in list(zip(bi, bo))[:150]:
        b1 = pm.Ball(radius=x[4], initPos=gm.Point(x[0], x[1]), initVel=gm.Point(x[2], x[3]))
        b2 = pm.Ball(radius=x[9], initPos=gm.Point(x[5], x[6]), initVel=gm.Point(x[7], x[8]))
        t, nr, pt = dy.get_toc_ball_ball(b1, b2, "ball-0", "ball-1")
        if not np.isfinite(ref[0]):
            assert nr is None and not np.isfinite(t)
            continue
        n_hit += 1
        # the device evaluates acos / asin / sin with the FMA-free portable routines (<= 2 ulp from libm)
        assert abs(t - ref[0]) <= 1e-12 * max(1.0, abs(ref[0]))
        assert abs(nr.x() - ref[1]) < 1e-9 and abs(nr.y() - ref[2]) < 1e-9
        assert abs(pt.x() - ref[3]) < 1e-8 and abs(pt.y() - ref[4]) < 1e-8
        fv = b1.futureVel_
        assert abs(fv.x() - ref[5]) < 1e-7 and abs(fv.y() - ref[6]) < 1e-7
    assert n_hit > 40


def _oracle_frame(model):
    case = None
    w = model.world_
    walls = [o for o in w.objects_.values() if not isinstance(o, pm.Ball)]
    balls = [o for o in w.objects_.values() if isinstance(o, pm.Ball)]
    names = list(w.objects_.keys())
    wi = {n: i for i, n in enumerate(k for k in names if k.startswith("wall"))}
    bi = {n: i for i, n in enumerate(k for k in names if k.startswith("ball"))}
    order = [wi[n] if n in wi else len(wi) + bi[n] for n in names]
    wv = np.array([[[v.x(), v.y()] for v in o.bbox_.vert_] for o in walls]).reshape(1, -1, 4, 2)
    pos = np.array([[[b.pos_.x(), b.pos_.y()] for b in balls]])
    rad = np.array([[b.get_radius() for b in balls]], np.float64)
    b0 = balls[0]
    return oracle.render_frames(wv, pos, rad, (int(w.xSz_), int(w.ySz_)), order=order,
                                wall_kind=[o.kind_ for o in walls], wall_rgba=np.array([o.draw_rgba() for o in walls]),
                                s_thick=int(b0.sThick_), fill=b0.fColor_.rgba(), stroke=b0.sColor_.rgba())[0]


def test_generate_image_matches_the_restated_renderer():
    model = mg.build_config1_64(NS)
    im0 = model.generate_image()
    assert im0.shape == (64, 64, 4) and im0.dtype == np.uint8
    assert np.array_equal(im0, _oracle_frame(model))
    for _ in range(30):
        model.step()
    im1 = model.generate_image()
    assert np.array_equal(im1, _oracle_frame(model))
    assert not np.array_equal(im0, im1)
    model = mg.build_cairo_trial_single(NS)      # WallDef walls, gray on white, 640x480
    model.step()
    im = model.world_.generate_image()
    assert im.shape == (480, 640, 4)
    assert np.array_equal(im, _oracle_frame(model))
    _, spr = pm.get_ball_im(radius=20, fColor=pm.Color(0.5, 0.5, 0.5), sThick=2)
    assert np.array_equal(spr, oracle.ball_sprite(20, 2))


def test_datasaver_save_writes_the_reference_tree(tmp_path):
    import scipy.io as sio
    from PIL import Image
    kw = dict(mg.RECT, numBalls=1, mnForce=8e4, mxForce=8e4, mnSeqLen=12, mxSeqLen=12, arenaSz=700)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=3792, svPrefix="val", **kw)
    ds.save(numSeq=3)
    for i in range(3):
        d = ds.seqDir_ % i
        assert sorted(os.listdir(d)) == ["data.mat"] + ["im%06d.jpg" % k for k in range(12)] + ["world.pkl"]
        mat = sio.loadmat(os.path.join(d, "data.mat"))
        assert mat["position"].shape == (2, 12) and mat["position"].dtype == np.float32
        assert mat["force"].shape == (2, 12) and np.count_nonzero(mat["force"][:, 1:]) == 0
        im = np.asarray(Image.open(os.path.join(d, "im000003.jpg")))
        assert im.shape == (700, 700, 3)
    # sequence 0 of seed 3792 is the golden savedata_nb1 world (same draws up to the sequence length)
    ref = CASES["savedata_nb1"]
    mat = sio.loadmat(os.path.join(ds.seqDir_ % 0, "data.mat"))
    assert np.array_equal(mat["force"][:, 0].astype(np.float64), ref["force"][0].astype(np.float32).astype(np.float64))
    import pickle
    with open(os.path.join(ds.seqDir_ % 0, "world.pkl"), "rb") as fh:
        wd = pickle.load(fh)
    assert sorted(wd.keys()) == ["ballPos", "force", "walls"]
    assert (wd["ballPos"][0].x(), wd["ballPos"][0].y()) == tuple(ref["ball_pos"][0])


def test_datasaver_batch_equals_single_world_stepping(tmp_path):
    """The sequences of one save()/fetch_batch() call run as one batch; each must equal the same
    sequence stepped alone through Dynamics.step() -- and the golden trajectory of the reference."""
    kw, seed = datasaver_kwargs("savedata_nb2")
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=seed, makeDirs=False, **kw)
    specs, position, frames = ds.fetch_batch(4, want_frames=False)
    ref = CASES["savedata_nb2"]
    got = position[0].T.reshape(-1, 2, 2)                      # (seqLen, balls, xy)
    assert np.array_equal(got, ref["traj"].astype(np.float32))
    ds2 = dio.DataSaver(rootPath=str(tmp_path), randSeed=seed, makeDirs=False, **kw)
    for k in range(2):
        ds2.seqLen_ = ds2._draw_seq_len()
        model, _, _, _ = ds2.generate_model()
        traj = run_model(model, 40)
        assert np.array_equal(traj.astype(np.float32), position[k].T.reshape(-1, 2, 2)[:40])


def test_datasaver_fetch_frames_and_crops(tmp_path):
    kw = dict(wThick=3, isRect=True, mnBallSz=3, mxBallSz=4, arenaSz=64, mnWLen=36, mxWLen=50, numBalls=2,
              mnForce=5.0, mxForce=12.0, mnSeqLen=20, mxSeqLen=30)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    ims = ds.fetch()
    assert len(ims) == ds.seqLen_ and ims[0].shape == (64, 64, 4)
    assert any(not np.array_equal(ims[0], im) for im in ims[1:])
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    imBalls, force, velocity, position = ds.fetch(cropSz=16, procSz=8)
    assert len(imBalls) == 2 and len(imBalls[0]) == ds.seqLen_ and imBalls[0][0].shape == (8, 8, 3)
    assert force.shape == velocity.shape == position.shape == (4, ds.seqLen_)
    # a crop is centred on its ball: the re-based position is the crop centre (up to rounding and clipping)
    assert np.all(np.abs(position - 4.0) <= 4.0)


def test_dynamics_horizon_delivers_delayed_frames():
    model = mg.build_config1_64(NS)
    hor = pm.DynamicsHorizon(model, lookAhead=5)
    im, mat = hor.get_data()
    assert im.shape == (64, 64, 4) and mat.shape == (1, 10) and mat.dtype == np.float32
    # constant velocity far from the walls: every displacement equals v * dt
    assert np.allclose(mat[0, 0::2], 60 * 0.01, atol=1e-6) and np.allclose(mat[0, 1::2], 25 * 0.01, atol=1e-6)


def test_dynamics_horizon_chunks_equal_single_steps():
    """The window is filled a chunk at a time on the device (Dynamics.step_many + ppb_horizon_async); frames and
    look-ahead matrices must be what stepping one step per get_data() gives."""
    N, calls = 4, 23
    a, b = mg.build_config1_64(NS), mg.build_config1_64(NS)
    for m in (a, b):
        m.set_object_velocity('ball-0', gm.Point(900.0, 410.0))      # fast: several wall hits inside the test
    hor = pm.DynamicsHorizon(a, lookAhead=N, chunk=6)
    p = b.get_object_position('ball-0')
    rows, ims = [np.array([p.x(), p.y()])], []
    for _ in range(N + calls):
        b.step()
        p = b.get_object_position('ball-0')
        rows.append(np.array([p.x(), p.y()]))
        ims.append(b.generate_image())
    for j in range(calls):
        im, mat = hor.get_data()
        assert np.array_equal(im, ims[j])
        exp = np.concatenate([np.float32(rows[j + k + 1] - rows[j + k]) for k in range(N)])
        assert np.array_equal(mat[0], exp), j


def test_datasaver_groups_large_jobs(tmp_path):
    """Jobs whose frames exceed the staging budget run as several batches with identical results."""
    kw = dict(wThick=3, isRect=True, mnBallSz=3, mxBallSz=3, arenaSz=64, mnWLen=36, mxWLen=50, numBalls=2,
              mnForce=5.0, mxForce=12.0, mnSeqLen=8, mxSeqLen=14)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=3, makeDirs=False, **kw)
    specs = ds.sample_sequences(11)
    p1, f1 = ds.run_sequences(specs, want_frames=True)
    p2, f2 = ds.run_sequences(specs, want_frames=True, max_stage_bytes=3 * 64 * 64 * 4)    # groups of 3 sequences
    seen = []
    p3, _ = ds.run_sequences(specs, want_frames=True, frame_sink=lambda s, i, im: seen.append((s, i, int(im.sum()))),
                             max_stage_bytes=4 * 64 * 64 * 4)
    for a, b, c in zip(p1, p2, p3):
        assert np.array_equal(a, b) and np.array_equal(a, c)
    for a, b in zip(f1, f2):
        assert np.array_equal(a, b)
    assert sorted(seen) == sorted((s, i, int(f1[s][i].sum())) for s in range(11) for i in range(specs[s].seq_len))
