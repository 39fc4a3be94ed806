"""Reference-facing modules (geometry / primitives / dataio) on the CPU: everything that happens
BEFORE the first step -- geometry helpers, cage construction, masses, apply_force, the DataSaver
random stream -- must be bit-identical with the unmodified reference.  The expected values are the
golden fixtures recorded from the reference (oracle/make_golden.py, oracle/make_toc_golden.py) and
the worlds are built with the SAME builder code that built them there, only with the façade
modules plugged in."""
import os
import pickle
import tempfile
import types

import numpy as np
import pytest

import make_golden as mg           # oracle/: world builders written against the reference API
from conftest import GOLDEN_DIR, load_golden

import pyphy_engine_b200.dataio as dio
import pyphy_engine_b200.geometry as gm
import pyphy_engine_b200.primitives as pm

NS = types.SimpleNamespace(gm=gm, pm=pm, dio=dio)
CASES = {c.name: c for c in load_golden("physics_golden.npz")}
TOC = np.load(os.path.join(GOLDEN_DIR, "toc_golden.npz"))


def extract(model):
    w = model.world_
    walls = [o for o in w.objects_.values() if not isinstance(o, pm.Ball)]
    balls = [o for o in w.objects_.values() if isinstance(o, pm.Ball)]
    names = list(w.objects_.keys())
    wi = {n: i for i, n in enumerate(k for k in names if k.startswith("wall"))}
    bi = {n: i for i, n in enumerate(k for k in names if k.startswith("ball"))}
    order = np.array([wi[n] if n in wi else len(wi) + bi[n] for n in names], np.int32)
    return dict(wall=np.array([[[v.x(), v.y()] for v in o.bbox_.vert_] for o in walls]).reshape(-1, 4, 2),
                pos0=np.array([[b.pos_.x(), b.pos_.y()] for b in balls]),
                vel0=np.array([[b.vel_.x(), b.vel_.y()] for b in balls]),
                mass=np.array([b.get_mass() for b in balls]),
                rad=np.array([b.get_radius() for b in balls], np.float64),
                order=order,
                wall_kind=np.array([o.kind_ for o in walls], np.int32))


HAND_BUILT = [("trial_single", mg.build_cairo_trial_single), ("trial_two", mg.build_cairo_trial_two),
              ("diamond", mg.build_diamond), ("config1_64", mg.build_config1_64),
              ("interleaved_g", mg.build_interleaved_order), ("freeze_block", mg.build_freeze_block)]


@pytest.mark.parametrize("name,builder", HAND_BUILT, ids=[n for n, _ in HAND_BUILT])
def test_hand_built_worlds_identical(name, builder):
    got, ref = extract(builder(NS)), CASES[name]
    for k, v in got.items():
        assert np.array_equal(v, ref[k]), k


def datasaver_kwargs(name):
    tail = name.rsplit("_s", 1)[-1]
    seed = int(tail) if tail.isdigit() else None
    if name.startswith("savedata_nb"):
        nb = int(name[-1])
        return dict(mg.RECT, numBalls=nb, mnForce=8e4, mxForce=8e4, mnSeqLen=210, mxSeqLen=210), \
            {1: 3792, 2: 4185, 3: 1094}[nb]
    if name.startswith("rect4"):
        return dict(mg.RECT, numBalls=4, mnForce=3e4, mxForce=8e4, mnSeqLen=150, mxSeqLen=150), seed
    if name.startswith("rect8"):
        return dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=1200, mnWLen=700, mxWLen=1000,
                    numBalls=8, mnForce=3e4, mxForce=8e4, mnSeqLen=120, mxSeqLen=120), seed
    if name.startswith("rhomb3"):
        return dict(wThick=30, isRect=False, wTheta=[30, 60], mnBallSz=15, mxBallSz=35, arenaSz=1600, mnWLen=500,
                    mxWLen=800, numBalls=3, mnForce=3e4, mxForce=8e4, mnSeqLen=150, mxSeqLen=150), seed
    if name.startswith("opp2"):
        return dict(mg.RECT, numBalls=2, mnForce=8e3, mxForce=5e4, mnSeqLen=120, mxSeqLen=120, oppForce=True), seed
    if name.startswith("rhombval"):
        return dict(wThick=20, isRect=False, wTheta=[23, 38, 45, 53], mnBallSz=15, mxBallSz=35, arenaSz=667,
                    mnWLen=200, mxWLen=300, numBalls=1, mnForce=1e3, mxForce=1e5, mnSeqLen=100, mxSeqLen=100), seed
    if name.startswith("crowd12"):
        return dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=900, mnWLen=420, mxWLen=520,
                    numBalls=12, mnForce=3e4, mxForce=8e4, mnSeqLen=200, mxSeqLen=200), seed
    return None, None


DS_CASES = [n for n in CASES if datasaver_kwargs(n)[0] is not None]


@pytest.mark.parametrize("name", DS_CASES)
def test_datasaver_random_stream_identical(name, tmp_path):
    """Same seed -> same sequence length, cage, ball centres, radii, forces and initial velocities
    as the reference's DataSaver (SURVEY.md Appendix D draw order)."""
    kw, seed = datasaver_kwargs(name)
    model, extra, seq_len = mg.build_datasaver(NS, seed, str(tmp_path), **kw)
    ref = CASES[name]
    assert seq_len == int(ref["seq_len"])
    for k in ("force", "ball_pos", "cage_pts"):
        assert np.array_equal(extra[k], ref[k]), k
    got = extract(model)
    for k, v in got.items():
        assert np.array_equal(v, ref[k]), k


def test_datasaver_second_sequence_continues_the_stream(tmp_path):
    """sample_sequences(n) draws exactly what n reference iterations draw: the first sequence equals
    the golden one and the generator state afterwards equals a hand-replayed stream."""
    kw, seed = datasaver_kwargs("rect4_s100")
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=seed, **kw)
    specs = ds.sample_sequences(3)
    ref = CASES["rect4_s100"]
    assert np.array_equal(np.array([[p.x(), p.y()] for p in specs[0].ball_pos]), ref["ball_pos"])
    ds2 = dio.DataSaver(rootPath=str(tmp_path), randSeed=seed, **kw)
    for k in range(3):
        ds2.seqLen_ = ds2._draw_seq_len()
        model, fs, bp, _ = ds2.generate_model()
        assert np.array_equal(np.array([[p.x(), p.y()] for p in bp]),
                              np.array([[p.x(), p.y()] for p in specs[k].ball_pos]))
    assert ds.rand_.rand() == ds2.rand_.rand()


def test_expstr_and_tree_names(tmp_path):
    ds = dio.DataSaver(rootPath=str(tmp_path), numBalls=2, mnBallSz=25, mxBallSz=25, mnSeqLen=210, mxSeqLen=210,
                       mnForce=8e4, mxForce=8e4, wThick=30, mnWLen=300, mxWLen=550, arenaSz=700, svPrefix="val")
    assert ds.expStr_ == "val-aSz700_wLen300-550_nb2_bSz25-25_f8.00e+04-8.00e+04_sLen210-210_wTh30"   # dataio.py:38-42
    assert os.path.isdir(os.path.join(str(tmp_path), ds.expStr_))
    assert ds.seqDir_ % 7 == os.path.join(str(tmp_path), ds.expStr_, "seq000007")
    ds = dio.DataSaver(rootPath=str(tmp_path), isRect=False, wTheta=[30, 60], oppForce=True, makeDirs=False)
    assert ds.expStr_.endswith("_wTh3030-_wTheta60_oppFrc")      # the reference's malformed join, dataio.py:44-52


# ------------------------------------------------------------------------------------------------
# geometry known answers (the reference's unit_tests.py:3-125, values recorded from the reference)
# ------------------------------------------------------------------------------------------------
def xy(p):
    return [np.nan, np.nan] if p is None else [p.x(), p.y()]


def same(a, b):
    return np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64), equal_nan=True)


def test_geometry_known_answers():
    P, L = gm.Point, gm.Line
    assert same([xy(L(P(-1, 0), P(1, 0)).get_intersection(L(P(0, -1), P(0, 1)))),
                 xy(L(P(0, 0), P(1, 1)).get_intersection(L(P(1, 0), P(0, 1)))),
                 xy(L(P(0, 0), P(1, 1)).get_intersection(L(P(1, 3), P(5, 7)))),
                 xy(L(P(0, -1), P(0, 1)).get_intersection(L(P(2, -1), P(2, 1))))], TOC["g_line_line"])
    l1 = L(P(-5, 5), P(5, 5))
    assert same([xy(l1.get_intersection_ray(L(P(-1, 0), P(1, 10)))), xy(l1.get_intersection_ray(L(P(-1, 0), P(-1, 3)))),
                 xy(l1.get_intersection_ray(L(P(-1, 3), P(-1, 0))))], TOC["g_line_ray"])
    bbox = gm.Bbox(P(1, 2), P(1, 1), P(2, 1), P(2, 2))
    lines = [L(P(0, 0), P(1.5, 1.9)), L(P(0, 0), P(1, 10)), L(P(0, 1), P(5, 1))]
    assert [bbox.is_intersect_line(l) for l in lines] == TOC["g_bbox_isect"].tolist() == [True, False, True]
    assert same([xy(p) + [d] for p, d in (bbox.get_intersection_with_line(l) for l in lines)], TOC["g_bbox_pt"])
    rays = [L(P(1.5, .5), P(1.5, -.5)), L(P(1.5, -.5), P(1.5, .5)), L(P(1.5, 1.9), P(0, 0))]
    assert [bbox.is_intersect_line_ray(l) for l in rays] == TOC["g_bbox_ray"].tolist() == [False, True, True]
    assert same([xy(P(1, 0).reflect_normal(P(-2, -1))), xy(P(0, 1).reflect_normal(P(-2, -1)))], TOC["g_reflect"])
    assert same(xy(L(P(0, 0), P(1, 1)).get_point_along_line(P(1, 1), 3)), TOC["g_along"])
    circ = gm.Circle(20, P(0, 0))
    tl = [L(P(-5, 25), P(5, 25)), L(P(5, 25), P(-5, 25)), L(P(-25, -5), P(-25, 5)), L(P(30, -5), P(30, 5)),
          L(P(50, 0), P(0, 50))]
    assert same([xy(circ.get_contact_point_pseudo_tangent(l)) for l in tl], TOC["g_tangent"])
    assert same([xy(L(P(0, 0), P(1, 1)).get_normal_towards_point(P(1, 0))),
                 xy(L(P(0, 0), P(1, 1)).get_normal_towards_point(P(0, 1)))], TOC["g_normal_towards"])


def test_point_semantics():
    p = gm.Point(3, 4)
    assert p.mag() == 5.0 and str(p) == "(3.00, 4.00)"
    q = p + 5            # Point + scalar: along its own direction (geometry.py:89-93)
    assert abs(q.x() - 6.0) < 1e-12 and abs(q.y() - 8.0) < 1e-12
    assert (p * 2).x() == 6.0 and (2 * p).y() == 8.0
    z = gm.Point(0, 0)
    z.make_unit_norm()
    assert (z.x(), z.y()) == (0.0, 0.0)
    assert gm.Point(2.5, -2.5).x_asint() == 3 and gm.Point(2.5, -2.5).y_asint() == -3     # py2 round
    assert gm.Point(0.5, 1.5).x_asint() == 1 and gm.Point(0.5, 1.5).y_asint() == 2
    l = gm.Line(gm.Point(0, 0), gm.Point(0, 5))      # |a| normalisation, geometry.py:203-216
    assert (l.a(), l.b(), l.c()) == (-1.0, 0.0, 0.0)
    l = gm.Line(gm.Point(0, 0), gm.Point(5, 0))      # a == 0: b and c stay unscaled
    assert (l.a(), l.b(), l.c()) == (0.0, 5.0, 0.0)
    with pytest.raises(AssertionError):
        gm.theta2dir(181)


def test_intersect_moving_circle_against_reference_vectors():
    """Circle.intersect_moving_circle (geometry.py:400-466) on the ball-ball golden inputs."""
    bi, bo = TOC["ball_in"], TOC["ball_out"]
    n_hit = 0
    for x, ref in zip(bi, bo):
        p1, v1, r1, p2, v2, r2 = gm.Point(x[0], x[1]), gm.Point(x[2], x[3]), x[4], gm.Point(x[5], x[6]), gm.Point(x[7], x[8]), x[9]
        rel = v2 - v1
        col = p2 - p1
        col.make_unit_norm()
        if -rel.dot(col) <= 0:
            assert not np.isfinite(ref[0])
            continue
        t, pt, nr = gm.Circle(r1, p1).intersect_moving_circle(gm.Circle(r2, p2), rel)
        if pt is None:
            assert not np.isfinite(ref[0])
            continue
        n_hit += 1
        assert t == ref[0] and (nr.x(), nr.y()) == (ref[1], ref[2]) and (pt.x(), pt.y()) == (ref[3], ref[4])
    assert n_hit > 100


def test_world_registry_semantics():
    world = pm.World(xSz=100, ySz=80)
    for w in pm.create_cage([gm.Point(5, 5), gm.Point(95, 5), gm.Point(95, 75), gm.Point(5, 75)], wThick=3):
        world.add_object(w)
    world.add_object(pm.WallDef(sz=gm.Point(10, 20)), initPos=gm.Point(40, 30))
    world.add_object(pm.BallDef(radius=4, density=7.0), initPos=gm.Point(20, 20))
    world.add_object(pm.Ball(radius=5), initPos=gm.Point(60, 60))
    assert world.get_object_names() == ["wall-0", "wall-1", "wall-2", "wall-3", "wall-4", "ball-0", "ball-1"]
    assert world.get_dynamic_object_names() == ["ball-0", "ball-1"]
    b0 = world.get_object("ball-0")
    assert b0.get_mass() == (4 / 3.0) * np.pi * np.power(4 / 10.0, 3)      # density dropped (primitives.py:394-400)
    p = b0.get_position()
    p.scale(2)
    assert b0.get_position().x() == 20.0                                    # getters copy
    b0.get_mutable_position().scale(2)
    assert b0.get_position().x() == 40.0                                    # mutable getters alias
    with pytest.raises(AssertionError):
        world.get_object("ball-9")
    with pytest.raises(Exception):
        world.add_object(object())
    model = pm.Dynamics(world)
    model.apply_force("ball-1", gm.Point(100.0, -50.0), forceT=1.0)         # vel += 0.5*T^2*F/m
    m = world.get_object("ball-1").get_mass()
    v = world.get_object("ball-1").get_velocity()
    assert (v.x(), v.y()) == ((0.5 * 1.0 * 1.0 * (100.0 * (1.0 / m))), (0.5 * 1.0 * 1.0 * (-50.0 * (1.0 / m))))
    world.del_object("ball-0")
    assert world.get_dynamic_object_names() == ["ball-1"]
    blob = pickle.dumps(world)
    w2 = pickle.loads(blob)
    assert w2.get_object_names() == world.get_object_names()
    assert w2.get_object("ball-1").get_position().x() == 60.0
