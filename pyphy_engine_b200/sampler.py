"""Vectorised host sampler for synthetic batches of random-initialised tables.

Draws worlds from the distribution of the reference's ``DataSaver`` rectangular
arena (``dataio.py:220-240`` walls, ``:309-379`` balls, ``:382-416`` forces) for
many worlds at once.  It is *distribution*-equivalent (exactly so for a fixed ball radius; with a radius range it redraws
the radius together with the position on every rejection, where the reference keeps the radius while
it searches a position -- the device sampler ``WorldBatch.sample_rect_worlds`` follows the reference's
loop structure), not RNG-stream-equivalent:
the stream-exact generator used by the drop-in ``DataSaver`` lives in
``pyphy_engine_b200/dataio.py``.  The cage geometry uses the same arithmetic as
``primitives.get_box_coords`` (``primitives.py:177-195``).
"""
from __future__ import annotations

import numpy as np


def _unit(v):
    """Point.make_unit_norm (geometry.py:48-51): multiply by 1/mag, zero stays zero."""
    mag = np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1])
    inv = np.where(mag > 0, 1.0 / np.where(mag > 0, mag, 1.0), 1.0)
    return np.stack([inv * v[..., 0], inv * v[..., 1]], -1)


def _line_abc(st, en):
    """Line.make_canonical (geometry.py:203-216)."""
    a = -(en[..., 1] - st[..., 1])
    b = en[..., 0] - st[..., 0]
    c = st[..., 0] * en[..., 1] - st[..., 1] * en[..., 0]
    am = np.abs(a)
    big = am > 1e-8
    d = np.where(big, am, 1.0)
    return np.where(big, a / d, 0.0), np.where(big, b / d, b), np.where(big, c / d, c)


def box_coords(st, en, w_thick):
    """get_box_coords (primitives.py:177-195) for arrays of segments -> (..., 4, 2)."""
    a, b, _ = _line_abc(st, en)
    nrml = _unit(np.stack([-a, -b], -1))
    diff = en - st
    dist = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])
    ldir = _unit(diff)
    pt1 = st
    pt2 = pt1 + ldir * dist[..., None]
    pt3 = pt2 - nrml * w_thick
    pt4 = pt3 - ldir * dist[..., None]
    return np.stack([pt1, pt2, pt3, pt4], -2)


def cage_vertices(pts, w_thick):
    """create_cage (primitives.py:199-209): pts (n, K, 2) -> walls (n, K, 4, 2)."""
    nxt = np.roll(pts, -1, axis=1)
    return box_coords(pts, nxt, float(w_thick))


def theta2dir(theta_deg):
    """geometry.theta2dir (geometry.py:177-188)."""
    th = np.pi * (np.asarray(theta_deg, np.float64) / 180.0)
    return _unit(np.stack([np.cos(th), np.sin(th)], -1))


def ball_mass(radius, density=1.0):
    """Ball.__init__ (primitives.py:390)."""
    return (4 / 3.0) * np.pi * np.power(np.asarray(radius, np.float64) / 10.0, 3) * density


def sample_rect_worlds(n_worlds, n_balls, seed=0, arena=700, w_thick=30, mn_wlen=300, mx_wlen=550,
                       mn_ball=25, mx_ball=25, mn_force=3e4, mx_force=8e4, force_t=1.0, max_tries=2000):
    """Returns dict(wall (n,4,4,2), pos (n,B,2), vel (n,B,2), rad (n,B), mass (n,B), cage_pts (n,4,2), force).

    Ball centres are integers at signed distance > r + w_thick + 2 from every cage
    line and pairwise non-overlapping (dataio.py:356-375); |F| = mn + floor(u*(mx-mn)),
    angle = +-u*pi with the same u (dataio.py:399-409); v = 0.5*forceT^2*F/m
    (primitives.py:580-594).
    """
    rng = np.random.RandomState(seed)
    n = int(n_worlds)
    T = float(w_thick)
    hlen = np.floor(mn_wlen + rng.rand(n) * (mx_wlen - mn_wlen))
    vlen = np.floor(mn_wlen + rng.rand(n) * (mx_wlen - mn_wlen))
    xleft = np.floor(rng.rand(n) * (arena - (hlen + T)))
    ytop = np.floor(rng.rand(n) * (arena - (vlen + T)))
    # _create_walls(xleft, ytop, (0, 90, 180), (hlen - T, vlen, hlen - T))  (dataio.py:228, 277-287)
    d1, d2, d3 = theta2dir(0.0), theta2dir(90.0), theta2dir(180.0)
    pt1 = np.stack([xleft, ytop], -1)
    pt2 = pt1 + d1 * (hlen - T)[:, None]
    pt3 = pt2 + d2 * vlen[:, None]
    pt4 = pt3 + d3 * (hlen - T)[:, None]
    pts = np.stack([pt1, pt2, pt3, pt4], 1)
    wall = cage_vertices(pts, T)
    la, lb, lc = _line_abc(pts, np.roll(pts, -1, axis=1))          # the four cage lines (dataio.py:296-297)
    ldn = np.sqrt(la * la + lb * lb)

    pos = np.zeros((n, n_balls, 2))
    rad = np.zeros((n, n_balls))
    for b in range(n_balls):
        todo = np.ones(n, bool)
        tries = 0
        while todo.any():
            idx = np.nonzero(todo)[0]
            m = idx.size
            r = np.floor(mn_ball + rng.rand(m) * (mx_ball - mn_ball))
            x = np.round(pts[idx, 0, 0] + rng.rand(m) * (pts[idx, 2, 0] - pts[idx, 0, 0]))
            y = np.round(pts[idx, 1, 1] + rng.rand(m) * (pts[idx, 3, 1] - pts[idx, 1, 1]))
            dist = (la[idx] * x[:, None] + lb[idx] * y[:, None] + lc[idx]) / ldn[idx]   # signed (geometry.py:244)
            ok = (dist > (r + T + 2)[:, None]).all(1)
            for j in range(b):
                dj = np.sqrt((x - pos[idx, j, 0]) ** 2 + (y - pos[idx, j, 1]) ** 2)
                ok &= dj > (rad[idx, j] + r)
            good = idx[ok]
            pos[good, b, 0], pos[good, b, 1], rad[good, b] = x[ok], y[ok], r[ok]
            todo[good] = False
            tries += 1
            if tries > max_tries:
                raise RuntimeError("could not place ball %d in %d worlds (cage too small?)" % (b, int(todo.sum())))
    mass = ball_mass(rad)
    u = rng.rand(n, n_balls)
    fmag = mn_force + np.floor(u * (mx_force - mn_force))
    ftheta = u * np.pi
    ftheta = np.where(rng.rand(n, n_balls) > 0.5, -ftheta, ftheta)
    force = np.stack([fmag * np.cos(ftheta), fmag * np.sin(ftheta)], -1)
    vel = (0.5 * force_t * force_t) * (force * (1.0 / mass)[..., None])
    return dict(wall=wall, pos=pos, vel=vel, rad=rad, mass=mass, cage_pts=pts, force=force)


def scaled64_params(n_balls):
    """The 64x64 variant of the rectangular arena (BASELINE configs 3/4): every
    length of the native 700-px setup scaled by 64/700 and rounded to integers >= 1, with the
    cage range widened so that `n_balls` balls fit; forces scaled so speeds shrink alike."""
    if n_balls <= 4:
        wl = (36, 50)
    else:
        wl = (50, 58)
    # native: F in [3e4, 8e4], m(r=25) = 65.45 -> v = 229..611 px/s; x 64/700 -> 21..56 px/s;
    # m(r=3) = 0.1131 -> F = 2 v m
    return dict(arena=64, w_thick=3, mn_wlen=wl[0], mx_wlen=wl[1], mn_ball=3, mx_ball=3,
                mn_force=4.74, mx_force=12.64)
