"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz from the reference.

Runs the UNMODIFIED reference (``/root/reference``) through ``oracle/ref_shim.py``
in the build container and records, for a set of worlds, everything a parity test
needs: initial state, wall vertices, per-step positions (fp64), the collision
event list (one record per ``Dynamics.resolve_collision`` call, in call order)
and the failure outcome when the reference raised.

    python oracle/make_golden.py            # rewrites tests/golden/

The reference tree is not present on the GPU box; the committed ``.npz`` files
are what travels.  Case lists are deterministic (fixed seeds).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# status codes shared with oracle/billiards_oracle.c and include/pyphy_b200.h
ST_OK, ST_TCOL_NEG, ST_OVERLAP, ST_THETA, ST_THETA_A, ST_INTPOINT, ST_AMISS, ST_DOMAIN, ST_TCOL_EQ = range(9)
_LINE_TO_STATUS = {
    ("primitives.py", 610): ST_TCOL_NEG,
    ("geometry.py", 413): ST_OVERLAP,
    ("geometry.py", 437): ST_THETA,
    ("geometry.py", 438): ST_THETA,
    ("geometry.py", 447): ST_THETA_A,
    ("dynamics.py", 34): ST_INTPOINT,
    ("dynamics.py", 45): ST_AMISS,
    ("geometry.py", 174): ST_DOMAIN,
    ("geometry.py", 444): ST_DOMAIN,
    ("primitives.py", 664): ST_TCOL_EQ,
}


def classify_exception(exc: BaseException) -> int:
    tb = traceback.extract_tb(exc.__traceback__)
    for fr in reversed(tb):
        key = (os.path.basename(fr.filename), fr.lineno)
        if key in _LINE_TO_STATUS:
            return _LINE_TO_STATUS[key]
    raise exc


def extract_world(ns, model):
    """Flatten a reference Dynamics/World into arrays (insertion order kept)."""
    pm = ns.pm
    world = model.world_
    walls, balls, order_names = [], [], []
    for name, obj in world.objects_.items():
        order_names.append(name)
        if isinstance(obj, pm.Ball):
            balls.append((name, obj))
        else:
            walls.append((name, obj))
    wall_idx = {n: i for i, (n, _) in enumerate(walls)}
    ball_idx = {n: i for i, (n, _) in enumerate(balls)}
    # dynamic order must equal ball insertion order
    assert list(world.dynamic_.keys()) == [n for n, _ in balls]
    nW = len(walls)
    order = np.array([wall_idx[n] if n in wall_idx else nW + ball_idx[n] for n in order_names], np.int32)
    wv = np.zeros((nW, 4, 2), np.float64)
    wkind = np.zeros((nW,), np.int32)
    for i, (_, w) in enumerate(walls):
        for k, v in enumerate(w.bbox_.vert_):
            wv[i, k] = (v.x(), v.y())
        wkind[i] = 1 if isinstance(w, pm.Wall) else 0
    pos = np.array([[b.pos_.x(), b.pos_.y()] for _, b in balls], np.float64).reshape(-1, 2)
    vel = np.array([[b.vel_.x(), b.vel_.y()] for _, b in balls], np.float64).reshape(-1, 2)
    rad = np.array([b.get_radius() for _, b in balls], np.float64)
    mass = np.array([b.get_mass() for _, b in balls], np.float64)
    name_to_obj = {}
    for n, i in wall_idx.items():
        name_to_obj[n] = i
    for n, i in ball_idx.items():
        name_to_obj[n] = nW + i
    return dict(order=order, wall=wv, wall_kind=wkind, pos0=pos, vel0=vel, rad=rad, mass=mass,
                dt=np.float64(model.deltaT_), g=np.array([model.g_.x(), model.g_.y()], np.float64),
                xsz=np.int32(world.xSz_), ysz=np.int32(world.ySz_)), ball_idx, name_to_obj, [b for _, b in balls]


def run_case(ns, model, n_steps):
    """Step the reference, logging trajectory, events, failure."""
    rec, ball_idx, name_to_obj, balls = extract_world(ns, model)
    nB = len(balls)
    traj = np.full((n_steps, nB, 2), np.nan, np.float64)
    events = []
    status, status_step = ST_OK, -1
    ns.events.clear()
    for s in range(n_steps):
        n0 = len(ns.events)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                model.step()
        except (AssertionError, ValueError, ZeroDivisionError, ref_shim.ReferenceTrap) as exc:
            status, status_step = classify_exception(exc), s
            for (_, name, partner) in ns.events[n0:]:
                events.append((s, ball_idx[name], name_to_obj[partner]))
            break
        for (_, name, partner) in ns.events[n0:]:
            events.append((s, ball_idx[name], name_to_obj[partner]))
        for j, b in enumerate(balls):
            traj[s, j] = (b.pos_.x(), b.pos_.y())
    velf = np.array([[b.vel_.x(), b.vel_.y()] for b in balls], np.float64).reshape(-1, 2)
    rec.update(traj=traj, vel_final=velf, events=np.array(events, np.int32).reshape(-1, 3),
               status=np.int32(status), status_step=np.int32(status_step), n_steps=np.int32(n_steps))
    return rec


# --------------------------------------------------------------------------- #
# world builders (all through the reference's own API)
# --------------------------------------------------------------------------- #
def build_cairo_trial_single(ns):
    """cairo_trial.create_single_ball_world_gray + get_horizon_data's force
    (cairo_trial.py:109-126, 211-214)."""
    gm, pm = ns.gm, ns.pm
    wThick = 30
    world = pm.World(xSz=640, ySz=480)
    bDef = pm.BallDef(fColor=pm.Color(0.5, 0.5, 0.5), radius=20)
    xLength, yLength = 550, 400
    wh = pm.WallDef(sz=gm.Point(xLength, wThick), fColor=pm.Color(0.5, 0.5, 0.5))
    wv = pm.WallDef(sz=gm.Point(wThick, yLength), fColor=pm.Color(0.5, 0.5, 0.5))
    xLeft, yTop = 30, 30
    world.add_object(wv, initPos=gm.Point(xLeft, yTop))
    world.add_object(wv, initPos=gm.Point(xLeft + xLength - wThick, yTop))
    world.add_object(wh, initPos=gm.Point(xLeft, yTop))
    world.add_object(wh, initPos=gm.Point(xLeft, yTop + yLength))
    world.add_object(bDef, initPos=gm.Point(200, 200))
    model = pm.Dynamics(world)
    model.apply_force('ball-0', gm.Point(-50000, 10000), forceT=1.0)
    return model


def build_cairo_trial_two(ns):
    """cairo_trial.create_multiple_ball_world_gray + multi_ball_world_simulation
    (cairo_trial.py:163-188): ball-0 launched at (1000, 0) into a resting ball."""
    gm, pm = ns.gm, ns.pm
    wThick = 30
    world = pm.World(xSz=640, ySz=480)
    b1 = pm.BallDef(fColor=pm.Color(0.5, 0.5, 0.5), radius=20)
    b2 = pm.BallDef(fColor=pm.Color(1.0, 1.0, 0.0), radius=20)
    xLength, yLength = 550, 400
    wh = pm.WallDef(sz=gm.Point(xLength, wThick), fColor=pm.Color(0.5, 0.5, 0.5))
    wv = pm.WallDef(sz=gm.Point(wThick, yLength), fColor=pm.Color(0.5, 0.5, 0.5))
    xLeft, yTop = 30, 30
    world.add_object(wv, initPos=gm.Point(xLeft, yTop))
    world.add_object(wv, initPos=gm.Point(xLeft + xLength - wThick, yTop))
    world.add_object(wh, initPos=gm.Point(xLeft, yTop))
    world.add_object(wh, initPos=gm.Point(xLeft, yTop + yLength))
    world.add_object(b1, initPos=gm.Point(200, 200))
    world.add_object(b2, initPos=gm.Point(400, 210))
    model = pm.Dynamics(world)
    model.world_.dynamic_['ball-0'].set_velocity(gm.Point(1000, 0))
    return model


def build_diamond(ns):
    """cairo_trial.create_world_diamond + ball_world_simulation (cairo_trial.py:61-75, 136-146)."""
    gm, pm = ns.gm, ns.pm
    world = pm.World(xSz=640, ySz=480)
    pts = [gm.Point(50, 240), gm.Point(300, 40), gm.Point(550, 240), gm.Point(300, 440)]
    for w in pm.create_cage(pts, wThick=20):
        world.add_object(w)
    world.add_object(pm.BallDef(fColor=pm.Color(0.5, 0.5, 0.5)), initPos=gm.Point(200, 200))
    model = pm.Dynamics(world)
    model.world_.dynamic_['ball-0'].set_velocity(gm.Point(2000, 100))
    return model


def build_ngon(ns, n_sides, n_balls, seed, radius=300.0, w_thick=20, r_ball=18):
    """BASELINE configs[4] 'polygonal wall layouts': a regular N-gon table through create_cage
    (primitives.py:199-209; vertices clockwise on screen so the walls thicken inwards), balls at
    non-overlapping random places inside the inscribed circle, random velocities."""
    gm, pm = ns.gm, ns.pm
    rs = np.random.RandomState(seed)
    cx = cy = 400.0
    world = pm.World(xSz=800, ySz=800)
    th0 = rs.rand() * 2 * np.pi
    pts = [gm.Point(cx + radius * np.cos(th0 + 2 * np.pi * i / n_sides), cy + radius * np.sin(th0 + 2 * np.pi * i / n_sides))
           for i in range(n_sides)]
    for w in pm.create_cage(pts, wThick=w_thick):
        world.add_object(w)
    free = radius * np.cos(np.pi / n_sides) - w_thick - r_ball - 2
    placed = []
    while len(placed) < n_balls:
        rr, aa = free * np.sqrt(rs.rand()), 2 * np.pi * rs.rand()
        x, y = np.floor(cx + rr * np.cos(aa)), np.floor(cy + rr * np.sin(aa))
        if all((x - px) ** 2 + (y - py) ** 2 > (2 * r_ball + 2) ** 2 for px, py in placed):
            placed.append((x, y))
    for x, y in placed:
        world.add_object(pm.BallDef(radius=r_ball, fColor=pm.Color(0.5, 0.5, 0.5)), initPos=gm.Point(x, y))
    model = pm.Dynamics(world)
    for j in range(n_balls):
        sp, aa = 300 + 900 * rs.rand(), 2 * np.pi * rs.rand()
        model.set_object_velocity('ball-%d' % j, gm.Point(sp * np.cos(aa), sp * np.sin(aa)))
    return model


def build_config1_64(ns):
    """BASELINE config 1: 64x64 world, 1 ball, 4-wall cage."""
    gm, pm = ns.gm, ns.pm
    world = pm.World(xSz=64, ySz=64)
    pts = [gm.Point(2, 2), gm.Point(62, 2), gm.Point(62, 62), gm.Point(2, 62)]
    for w in pm.create_cage(pts, wThick=3, fColor=pm.Color(1.0, 0.0, 0.0)):
        world.add_object(w)
    world.add_object(pm.BallDef(radius=4, sThick=1, fColor=pm.Color(0.5, 0.5, 0.5)), initPos=gm.Point(20, 30))
    model = pm.Dynamics(world, g=0, deltaT=0.01)
    model.set_object_velocity('ball-0', gm.Point(60, 25))
    return model


def build_interleaved_order(ns):
    """Walls and balls added in interleaved order and gravity switched on:
    exercises the insertion-order tie-break (primitives.py:724-733) and the
    g terms of move_object (primitives.py:604-606)."""
    gm, pm = ns.gm, ns.pm
    world = pm.World(xSz=400, ySz=400)
    cage = pm.create_cage([gm.Point(20, 20), gm.Point(380, 20), gm.Point(380, 380), gm.Point(20, 380)], wThick=10)
    world.add_object(cage[0])
    world.add_object(pm.BallDef(radius=15), initPos=gm.Point(100, 100))
    world.add_object(cage[1])
    world.add_object(pm.BallDef(radius=22), initPos=gm.Point(250, 120))
    world.add_object(cage[2])
    world.add_object(cage[3])
    world.add_object(pm.BallDef(radius=18), initPos=gm.Point(200, 300))
    model = pm.Dynamics(world, g=300.0, deltaT=0.01)
    model.set_object_velocity('ball-0', gm.Point(900, 200))
    model.set_object_velocity('ball-1', gm.Point(-700, 350))
    model.set_object_velocity('ball-2', gm.Point(100, -800))
    return model


def build_freeze_block(ns):
    """A ball that enters a free-standing block through its corner (corner contacts are
    not detected, TODO:6-10), hits a face from the inside and ends in the frozen state
    of primitives.py:647-650 (computed t <= 0 with an empty queue)."""
    gm, pm = ns.gm, ns.pm
    world = pm.World(xSz=400, ySz=400)
    for w in pm.create_cage([gm.Point(20, 20), gm.Point(380, 20), gm.Point(380, 380), gm.Point(20, 380)], wThick=10):
        world.add_object(w)
    world.add_object(pm.WallDef(sz=gm.Point(100, 100)), initPos=gm.Point(150, 150))
    world.add_object(pm.BallDef(radius=20), initPos=gm.Point(100, 100))
    world.add_object(pm.BallDef(radius=5), initPos=gm.Point(330, 60))
    model = pm.Dynamics(world, g=0, deltaT=0.01)
    model.world_.dynamic_["ball-0"].set_velocity(gm.Point(300, 300))
    model.world_.dynamic_["ball-1"].set_velocity(gm.Point(0, -900))
    return model


def build_datasaver(ns, seed, tmpdir, **kw):
    """One sequence exactly as DataSaver.save draws it (dataio.py:88-99, 213-217)."""
    dio = ns.dio
    with contextlib.redirect_stdout(io.StringIO()):
        ds = dio.DataSaver(rootPath=tmpdir, randSeed=seed, **kw)
        seqLen = int(ds.mnSeqLen_ + ds.rand_.rand() * (ds.mxSeqLen_ - ds.mnSeqLen_))
        ds.seqLen_ = seqLen
        model, f, ballPos, walls = ds.generate_model()
    extra = dict(force=np.array([[p.x(), p.y()] for p in f], np.float64),
                 ball_pos=np.array([[p.x(), p.y()] for p in ballPos], np.float64),
                 cage_pts=np.array([[p.x(), p.y()] for p in ds.pts], np.float64),
                 seq_len=np.int32(seqLen), seed=np.int32(seed))
    return model, extra, seqLen


def build_failure(ns, kind):
    """Hand-built degenerate worlds on which the reference itself raises; they pin
    the per-world status codes (failure detection)."""
    gm, pm = ns.gm, ns.pm
    world = pm.World(xSz=400, ySz=400)
    for w in pm.create_cage([gm.Point(20, 20), gm.Point(380, 20), gm.Point(380, 380), gm.Point(20, 380)], wThick=10):
        world.add_object(w)
    if kind == "overlap":          # geometry.py:413 -- balls interpenetrate by more than 0.1
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(100, 200))
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(130, 200))
        vels = {"ball-0": (100, 0), "ball-1": (-100, 0)}
    elif kind == "overlap_late":   # the same assertion, reached after some ordinary steps
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(100, 200))
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(300, 200))
        world.add_object(pm.BallDef(radius=5), initPos=gm.Point(200, 100))
        vels = {"ball-0": (500, 0), "ball-1": (-500, 0), "ball-2": (0, 100)}
    elif kind == "centre_on_line":  # dynamics.py:34 -- ball centre on the (extended) line of an edge
        world.add_object(pm.BallDef(radius=3), initPos=gm.Point(200, 30))
        vels = {"ball-0": (50, -80)}
    elif kind == "amiss":          # dynamics.py:43-45 -- ball already straddles an edge and moves into it
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(45, 200))
        vels = {"ball-0": (-100, 0)}
    elif kind == "corner_clip":    # TODO:6-10 -- corner contacts are not detected; the ball enters a
        # block through its corner and the next rescan (triggered by ball-1) trips dynamics.py:43-45
        world.add_object(pm.WallDef(sz=gm.Point(100, 100)), initPos=gm.Point(150, 150))
        world.add_object(pm.BallDef(radius=20), initPos=gm.Point(100, 110))
        world.add_object(pm.BallDef(radius=5), initPos=gm.Point(330, 60))
        vels = {"ball-0": (300, 200), "ball-1": (0, -9000)}
    else:
        raise ValueError(kind)
    model = pm.Dynamics(world, g=0, deltaT=0.01)
    for name, v in vels.items():
        model.world_.dynamic_[name].set_velocity(gm.Point(*v))
    if kind == "overlap_late":
        # teleport ball-1 into ball-0 after the caches were built: set_position does not
        # refresh tCol_ (primitives.py:462), the overlap is met at the next rescan
        for _ in range(3):
            model.step()
        model.world_.dynamic_["ball-1"].set_position(model.world_.dynamic_["ball-0"].get_position() + gm.Point(25, 0))
        model.add_all_dynamic_collision_queue()
    return model


RECT = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=700, mnWLen=300, mxWLen=550)


def main_failures(ns):
    out, cases = {}, []
    for kind in ("overlap", "overlap_late", "centre_on_line", "amiss", "corner_clip"):
        rec = run_case(ns, build_failure(ns, kind), 40)
        assert int(rec["status"]) != ST_OK, kind
        for k, v in rec.items():
            out["fail_%s/%s" % (kind, k)] = v
        cases.append("fail_" + kind)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "failure_golden.npz"), **out)
    print("failure cases:", [(c, int(out[c + "/status"]), int(out[c + "/status_step"])) for c in cases])


def main():
    ns = ref_shim.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    main_failures(ns)
    if "--only-failures" in sys.argv:
        return
    tmpdir = tempfile.mkdtemp(prefix="pyphy_golden_")
    out = {}

    def put(prefix, rec, extra=None):
        for k, v in rec.items():
            out["%s/%s" % (prefix, k)] = v
        for k, v in (extra or {}).items():
            out["%s/%s" % (prefix, k)] = v

    cases = []
    # hand-built worlds from the reference's own demos
    put("trial_single", run_case(ns, build_cairo_trial_single(ns), 100)); cases.append("trial_single")
    put("trial_two", run_case(ns, build_cairo_trial_two(ns), 300)); cases.append("trial_two")
    put("diamond", run_case(ns, build_diamond(ns), 100)); cases.append("diamond")
    put("config1_64", run_case(ns, build_config1_64(ns), 100)); cases.append("config1_64")
    put("interleaved_g", run_case(ns, build_interleaved_order(ns), 200)); cases.append("interleaved_g")
    put("freeze_block", run_case(ns, build_freeze_block(ns), 80)); cases.append("freeze_block")
    # save_data.py's three validation configs (save_data.py:3-12), first sequence of each
    for nb, seed in zip((1, 2, 3), (3792, 4185, 1094)):
        kw = dict(RECT, numBalls=nb, mnForce=8e4, mxForce=8e4, mnSeqLen=210, mxSeqLen=210)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "savedata_nb%d" % nb
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # BASELINE config 2: 4 balls, ICLR16 rect arena (dataio.py:435-446 parameters)
    for seed in range(100, 124):
        kw = dict(RECT, numBalls=4, mnForce=3e4, mxForce=8e4, mnSeqLen=150, mxSeqLen=150)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "rect4_s%d" % seed
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # 8 balls, larger arena (config 4's ball count at native scale)
    for seed in range(200, 208):
        kw = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=1200, mnWLen=700, mxWLen=1000,
                  numBalls=8, mnForce=3e4, mxForce=8e4, mnSeqLen=120, mxSeqLen=120)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "rect8_s%d" % seed
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # non-rectangular (rhombus) tables, mixed ball sizes (dataio.py:243-306; parameter
    # ranges of save_nonrect_arena_train, dataio.py:424-432)
    for seed in range(300, 310):
        kw = dict(wThick=30, isRect=False, wTheta=[30, 60], mnBallSz=15, mxBallSz=35, arenaSz=1600,
                  mnWLen=500, mxWLen=800, numBalls=3, mnForce=3e4, mxForce=8e4, mnSeqLen=150, mxSeqLen=150)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "rhomb3_s%d" % seed
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # oppForce=True: the two balls are kicked towards each other (dataio.py:386-397)
    for seed in range(600, 604):
        kw = dict(RECT, numBalls=2, mnForce=8e3, mxForce=5e4, mnSeqLen=120, mxSeqLen=120, oppForce=True)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "opp2_s%d" % seed
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # the validation rhombus set: four angles (dataio.py:419-422 parameters)
    for seed in range(700, 704):
        kw = dict(wThick=20, isRect=False, wTheta=[23, 38, 45, 53], mnBallSz=15, mxBallSz=35, arenaSz=667,
                  mnWLen=200, mxWLen=300, numBalls=1, mnForce=1e3, mxForce=1e5, mnSeqLen=100, mxSeqLen=100)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        name = "rhombval_s%d" % seed
        put(name, run_case(ns, model, seqLen), extra); cases.append(name)
    # regular N-gon tables through create_cage (BASELINE configs[4]: polygonal wall layouts)
    for n_sides, n_balls in ((3, 2), (5, 3), (6, 4), (8, 6)):
        for seed in (800, 801):
            name = "ngon%d_s%d" % (n_sides, seed)
            put(name, run_case(ns, build_ngon(ns, n_sides, n_balls, seed), 150)); cases.append(name)
    # crowded worlds: search for seeds on which the reference itself raises, to pin
    # the failure outcomes (status code + step)
    n_fail = 0
    for seed in range(400, 520):
        kw = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=900, mnWLen=420, mxWLen=520,
                  numBalls=12, mnForce=3e4, mxForce=8e4, mnSeqLen=200, mxSeqLen=200)
        model, extra, seqLen = build_datasaver(ns, seed, tmpdir, **kw)
        rec = run_case(ns, model, seqLen)
        keep = False
        if int(rec["status"]) != ST_OK and n_fail < 8:
            n_fail += 1
            keep = True
        elif seed < 404:
            keep = True
        if keep:
            name = "crowd12_s%d" % seed
            put(name, rec, extra); cases.append(name)
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "physics_golden.npz"), **out)
    nev = sum(len(out[c + "/events"]) for c in cases)
    nfail = sum(int(out[c + "/status"]) != 0 for c in cases)
    print("wrote %d cases, %d events, %d reference failures" % (len(cases), nev, nfail))


if __name__ == "__main__":
    main()
