"""Time-of-collision helpers with the reference's ``dynamics.py`` interface.

``get_toc_ball_wall(ball, wall)`` and ``get_toc_ball_ball(ball1, ball2, name1, name2)`` keep the
reference's signatures and return conventions ``(tCol, nrmlCol, ptCol)`` -- ``(inf, None, None)``
when the objects never meet (dynamics.py:8-63, 67-136).  The arithmetic is not done here: the
pair is uploaded as a one-world batch and evaluated by the scan stage of the advance kernel
(``ppb_scan_async``, csrc/ppb_kernels.cuh + ppb_math.cuh ``Compat``), i.e. by exactly the code
that ``Dynamics.step`` uses, in the reference's fp64 operation order.

``ptCol`` is informational in the reference (stored in ``Dynamics.ptCol_`` and never read); it is
reconstructed here from the time and normal the kernel returns:
  * wall: the point of the wall line that the ball touches, ``pos + t*vel - r*nrmlCol``
    (the reference's ``intPoint + t*velOrth``, dynamics.py:50);
  * ball: the centre of ball 2 at contact in the frame in which ball 1 rests,
    ``pos2 + t*(vel2 - vel1)`` (``newP2 + p1``, geometry.py:459-465).
As in the reference, ``get_toc_ball_ball`` stores ball 1's post-collision velocity in
``ball1.futureVel_`` as a side effect (dynamics.py:129).
"""
from __future__ import annotations

import numpy as np

from . import geometry as gm

__all__ = ["get_toc_ball_wall", "get_toc_ball_ball"]


def _status_error(status):
    from ._capi import STATUS_NAMES
    return AssertionError("the reference raises here: %s" % STATUS_NAMES.get(int(status), status))


def _scan_pair(balls, walls, device=0):
    """Upload (balls, walls) as one world, run the TOC scan for every ball, return the caches."""
    from .engine import WorldBatch
    nB, nW = len(balls), len(walls)
    wb = WorldBatch(1, nB, nW, dtype="f64", device=device)
    try:
        wv = np.zeros((1, nW, 4, 2), np.float64)
        for i, w in enumerate(walls):
            for k, l in enumerate(w.get_lines()):
                st = l.st()
                wv[0, i, k] = (st.x(), st.y())
        wb.set_walls(wv)
        pos = np.array([[[b.get_position().x(), b.get_position().y()] for b in balls]], np.float64)
        vel = np.array([[[b.get_velocity().x(), b.get_velocity().y()] for b in balls]], np.float64)
        rad = np.array([[b.get_radius() for b in balls]], np.float64)
        mass = np.array([[b.get_mass() for b in balls]], np.float64)
        wb.set_balls(pos, vel, rad, mass)
        wb.scan()
        status, _ = wb.get_status()
        if status[0] != 0:
            raise _status_error(status[0])
        tcol, partner = wb.get_collision_cache()
        nrml, fv = wb.get_collision_response()
        return tcol[0], partner[0], nrml[0], fv[0]
    finally:
        wb.close()


def get_toc_ball_wall(obj1, obj2):
    """obj1: ball, obj2: wall (``Wall`` or ``GenericWall``) -> (tCol, nrmlCol, ptCol)."""
    tcol, partner, nrml, _ = _scan_pair([obj1], [obj2])
    if partner[0] < 0 or not np.isfinite(tcol[0]):
        return np.inf, None, None
    t = float(tcol[0])
    n = gm.Point(nrml[0, 0], nrml[0, 1])
    pos, vel, r = obj1.get_position(), obj1.get_velocity(), obj1.get_radius()
    pt = pos + (t * vel) - (r * n)
    return t, n, pt


def get_toc_ball_ball(obj1, obj2, name1=None, name2=None):
    """-> (tCol, contact tangent, centre of ball 2 at contact); sets obj1.futureVel_."""
    tcol, partner, nrml, fv = _scan_pair([obj1, obj2], [])
    if partner[0] < 0 or not np.isfinite(tcol[0]):
        return np.inf, None, None
    t = float(tcol[0])
    n = gm.Point(nrml[0, 0], nrml[0, 1])
    rel = obj2.get_velocity() - obj1.get_velocity()
    pt = obj2.get_position() + (t * rel)
    obj1.set_after_collision_velocity(gm.Point(fv[0, 0], fv[0, 1]))
    return t, n, pt
