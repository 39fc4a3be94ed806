"""TEST INFRASTRUCTURE ONLY -- golden vectors for the reference's time-of-collision functions.

Runs the UNMODIFIED ``dynamics.get_toc_ball_wall`` / ``get_toc_ball_ball`` (dynamics.py:8-136) and
the geometry known-answer calls of ``unit_tests.py`` through ``oracle/ref_shim.py`` and writes
``tests/golden/toc_golden.npz``:

  wall cases : ball (pos, vel, r) x GenericWall(st, en, wThick)  -> t, nrml, pt
  ball cases : two balls (pos, vel, r)                            -> t, tangent, pt, futureVel of ball 1
  geometry   : the unit_tests.py calls with the values the reference returns

    python oracle/make_toc_golden.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ns = ref_shim.load()
    gm, pm, dy = ns.gm, ns.pm, ns.dy
    rng = np.random.RandomState(77)
    wall_in, wall_out = [], []
    for _ in range(400):
        st = gm.Point(*rng.uniform(50, 600, 2))
        ang = rng.uniform(-np.pi, np.pi)
        length = rng.uniform(40, 400)
        if rng.rand() < 0.4:   # axis-aligned walls are the common case
            ang = rng.choice([0.0, 0.5 * np.pi, np.pi, -0.5 * np.pi])
        en = gm.Point(st.x() + length * np.cos(ang), st.y() + length * np.sin(ang))
        th = float(rng.choice([3, 10, 20, 30]))
        wall = pm.GenericWall(st, en, wThick=th)
        r = int(rng.choice([3, 15, 20, 25]))
        # a ball somewhere off the wall
        mid = gm.Point(0.5 * (st.x() + en.x()), 0.5 * (st.y() + en.y()))
        off = rng.uniform(r + th + 5, 300)
        side = rng.choice([-1.0, 1.0])
        n = wall.bbox_.get_lines()[0].get_normal()
        pos = gm.Point(mid.x() + side * off * n.x() + rng.uniform(-60, 60), mid.y() + side * off * n.y() + rng.uniform(-60, 60))
        vel = gm.Point(*rng.uniform(-900, 900, 2))
        ball = pm.Ball(radius=r, initPos=pos, initVel=vel)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                t, nr, pt = dy.get_toc_ball_wall(ball, wall)
        except (AssertionError, ref_shim.ReferenceTrap):
            continue
        wall_in.append([st.x(), st.y(), en.x(), en.y(), th, pos.x(), pos.y(), vel.x(), vel.y(), r])
        if nr is None:
            wall_out.append([np.inf, 0, 0, 0, 0])
        else:
            wall_out.append([t, nr.x(), nr.y(), pt.x(), pt.y()])
    ball_in, ball_out = [], []
    for _ in range(400):
        r1, r2 = int(rng.choice([3, 15, 25, 35])), int(rng.choice([3, 15, 25, 35]))
        p1 = gm.Point(*rng.uniform(100, 600, 2))
        d = rng.uniform(r1 + r2 + 0.5, 400)
        a = rng.uniform(-np.pi, np.pi)
        if rng.rand() < 0.1:
            d = r1 + r2 - rng.uniform(0, 0.09)      # inside the [R-0.1, R] nudge band (geometry.py:411)
        p2 = gm.Point(p1.x() + d * np.cos(a), p1.y() + d * np.sin(a))
        v1 = gm.Point(*rng.uniform(-600, 600, 2))
        v2 = gm.Point(*rng.uniform(-600, 600, 2))
        if rng.rand() < 0.5:   # aim ball 2 roughly at ball 1 so that many cases do collide
            s = rng.uniform(100, 900)
            v2 = gm.Point(v1.x() - s * np.cos(a + rng.uniform(-0.3, 0.3)), v1.y() - s * np.sin(a + rng.uniform(-0.3, 0.3)))
        b1 = pm.Ball(radius=r1, initPos=p1, initVel=v1)
        b2 = pm.Ball(radius=r2, initPos=p2, initVel=v2)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                t, nr, pt = dy.get_toc_ball_ball(b1, b2, "ball-0", "ball-1")
        except (AssertionError, ValueError, ref_shim.ReferenceTrap):
            continue
        ball_in.append([p1.x(), p1.y(), v1.x(), v1.y(), r1, p2.x(), p2.y(), v2.x(), v2.y(), r2])
        if nr is None:
            ball_out.append([np.inf, 0, 0, 0, 0, 0, 0])
        else:
            ball_out.append([t, nr.x(), nr.y(), pt.x(), pt.y(), b1.futureVel_.x(), b1.futureVel_.y()])
    out = dict(wall_in=np.array(wall_in), wall_out=np.array(wall_out), ball_in=np.array(ball_in),
               ball_out=np.array(ball_out))

    # ---- geometry known answers (unit_tests.py:3-125), evaluated by the reference itself ----
    P, L = gm.Point, gm.Line
    def xy(p):
        return [np.nan, np.nan] if p is None else [p.x(), p.y()]
    out["g_line_line"] = np.array([
        xy(L(P(-1, 0), P(1, 0)).get_intersection(L(P(0, -1), P(0, 1)))),
        xy(L(P(0, 0), P(1, 1)).get_intersection(L(P(1, 0), P(0, 1)))),
        xy(L(P(0, 0), P(1, 1)).get_intersection(L(P(1, 3), P(5, 7)))),
        xy(L(P(0, -1), P(0, 1)).get_intersection(L(P(2, -1), P(2, 1))))])
    l1 = L(P(-5, 5), P(5, 5))
    out["g_line_ray"] = np.array([xy(l1.get_intersection_ray(L(P(-1, 0), P(1, 10)))),
                                  xy(l1.get_intersection_ray(L(P(-1, 0), P(-1, 3)))),
                                  xy(l1.get_intersection_ray(L(P(-1, 3), P(-1, 0))))])
    bbox = gm.Bbox(P(1, 2), P(1, 1), P(2, 1), P(2, 2))
    lines = [L(P(0, 0), P(1.5, 1.9)), L(P(0, 0), P(1, 10)), L(P(0, 1), P(5, 1))]
    out["g_bbox_isect"] = np.array([bbox.is_intersect_line(l) for l in lines])
    res = [bbox.get_intersection_with_line(l) for l in lines]
    out["g_bbox_pt"] = np.array([xy(p) + [d] for p, d in res])
    rays = [L(P(1.5, .5), P(1.5, -.5)), L(P(1.5, -.5), P(1.5, .5)), L(P(1.5, 1.9), P(0, 0))]
    out["g_bbox_ray"] = np.array([bbox.is_intersect_line_ray(l) for l in rays])
    out["g_reflect"] = np.array([xy(P(1, 0).reflect_normal(P(-2, -1))), xy(P(0, 1).reflect_normal(P(-2, -1)))])
    out["g_along"] = np.array(xy(L(P(0, 0), P(1, 1)).get_point_along_line(P(1, 1), 3)))
    circ = gm.Circle(20, P(0, 0))
    tl = [L(P(-5, 25), P(5, 25)), L(P(5, 25), P(-5, 25)), L(P(-25, -5), P(-25, 5)), L(P(30, -5), P(30, 5)),
          L(P(50, 0), P(0, 50))]
    out["g_tangent"] = np.array([xy(circ.get_contact_point_pseudo_tangent(l)) for l in tl])
    out["g_normal_towards"] = np.array([xy(L(P(0, 0), P(1, 1)).get_normal_towards_point(P(1, 0))),
                                        xy(L(P(0, 0), P(1, 1)).get_normal_towards_point(P(0, 1)))])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "toc_golden.npz"), **out)
    print("wall cases %d (%d hit), ball cases %d (%d hit)" % (
        len(wall_in), int(np.isfinite(out["wall_out"][:, 0]).sum()), len(ball_in),
        int(np.isfinite(out["ball_out"][:, 0]).sum())))
    for k in sorted(out):
        if k.startswith("g_"):
            print(k, out[k].tolist())


if __name__ == "__main__":
    main()
