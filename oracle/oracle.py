"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/billiards_oracle.c.

Importable from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs; never from the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "_oracle.so")
SOURCES = [os.path.join(HERE, "billiards_oracle.c"), os.path.join(HERE, "render_oracle.c"),
           os.path.join(HERE, "cairo_oracle.c")]

_c_dp = ctypes.POINTER(ctypes.c_double)
_c_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (no FMA contraction, IEEE fp64)."""
    srcs = [s for s in SOURCES if os.path.exists(s)]
    hdr = os.path.join(os.path.dirname(HERE), "pyphy_engine_b200", "csrc", "portable_trig.h")
    newest = max(os.path.getmtime(p) for p in srcs + [hdr])
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < newest:
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
               "-Wall", "-o", SO_PATH] + srcs + ["-lm"]
        subprocess.check_call(cmd)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(SO_PATH)
        _lib.oracle_step_batch.restype = ctypes.c_int64
        _lib.oracle_trig.restype = ctypes.c_double
        _lib.oracle_trig.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int]
        _lib.oracle_disc_box_area.restype = ctypes.c_double
        _lib.oracle_quad_pixel_area.restype = ctypes.c_double
        _lib.cairo_arc_segments_needed.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double]
    return _lib


def _dp(a):
    return a.ctypes.data_as(_c_dp)


def _ip(a):
    return a.ctypes.data_as(_c_ip)


class OracleBatch:
    """State of ``n_worlds`` independent worlds with a common object layout,
    stepped by the C restatement of ``Dynamics.step`` (primitives.py:628-656).

    wall: (n_worlds, n_walls, 4, 2) vertices; pos/vel: (n_worlds, n_balls, 2);
    rad/mass: (n_worlds, n_balls) (rad < 0 marks an empty slot); order: object
    insertion order, walls numbered 0..n_walls-1 and balls n_walls..; g = (gx, gy).
    """

    def __init__(self, wall, pos, vel, rad, mass, order=None, dt=0.01, g=(0.0, 0.0), trig_mode=0,
                 max_substeps=0, friction=0.0):
        f8 = np.float64
        self.wall = np.ascontiguousarray(wall, f8)
        self.pos = np.ascontiguousarray(pos, f8).copy()
        self.vel = np.ascontiguousarray(vel, f8).copy()
        self.n_worlds, self.n_balls = self.pos.shape[:2]
        self.n_walls = self.wall.shape[1]
        assert self.wall.shape == (self.n_worlds, self.n_walls, 4, 2)
        self.rad = np.ascontiguousarray(np.broadcast_to(rad, (self.n_worlds, self.n_balls)), f8).copy()
        self.mass = np.ascontiguousarray(np.broadcast_to(mass, (self.n_worlds, self.n_balls)), f8).copy()
        if order is None:
            order = np.arange(self.n_walls + self.n_balls)
        self.order = np.ascontiguousarray(order, np.int32)
        assert self.order.shape == (self.n_walls + self.n_balls,)
        self.dt, self.g, self.trig_mode = float(dt), (float(g[0]), float(g[1])), int(trig_mode)
        self.max_substeps = int(max_substeps)
        self.friction = float(friction)     # engine extension, 0 = the reference (see billiards_oracle.c)
        nw, nb = self.n_worlds, self.n_balls
        # Dynamics._include_object, primitives.py:556-563
        self.tcol = np.zeros((nw, nb), f8)
        self.fvel = np.zeros((nw, nb, 2), f8)
        self.nrml = np.zeros((nw, nb, 2), f8)
        self.partner = np.full((nw, nb), -1, np.int32)
        self.pending = np.ones((nw,), np.int32)
        self.status = np.zeros((nw,), np.int32)
        self.status_step = np.full((nw,), -1, np.int32)
        self.steps_done = 0
        self.n_events_total = 0

    def step(self, n_steps=1, record_traj=False, max_events=0, n_threads=1):
        """Advance every world by ``n_steps`` calls of ``Dynamics.step``.

        Returns (traj or None, events or None); traj is (n_steps, n_worlds,
        n_balls, 2); events is (k, 4) int32 rows (world, step, ball, partner).
        """
        L = lib()
        L.oracle_set_friction(ctypes.c_double(self.friction))
        traj = np.empty((n_steps, self.n_worlds, self.n_balls, 2), np.float64) if record_traj else None
        ev = np.zeros((max_events, 4), np.int32) if max_events > 0 else None
        n = L.oracle_step_batch(
            ctypes.c_int(self.n_worlds), ctypes.c_int(self.n_balls), ctypes.c_int(self.n_walls),
            _ip(self.order), _dp(self.pos), _dp(self.vel), _dp(self.rad), _dp(self.mass), _dp(self.wall),
            _dp(self.tcol), _dp(self.fvel), _dp(self.nrml), _ip(self.partner), _ip(self.pending),
            _ip(self.status), _ip(self.status_step),
            ctypes.c_double(self.dt), ctypes.c_double(self.g[0]), ctypes.c_double(self.g[1]),
            ctypes.c_int(self.steps_done), ctypes.c_int(n_steps),
            _dp(traj) if traj is not None else None,
            _ip(ev) if ev is not None else None, ctypes.c_int64(max_events),
            ctypes.c_int(self.trig_mode), ctypes.c_int(n_threads), ctypes.c_int(self.max_substeps))
        self.steps_done += n_steps
        self.n_events_total += int(n)
        if ev is not None:
            if n > max_events:
                raise RuntimeError("event buffer too small: %d > %d" % (n, max_events))
            ev = ev[:n]
        return traj, ev


def render_frames(wall, pos, rad, canvas, order=None, wall_kind=None, wall_rgba=None, s_thick=2,
                  fill=(0.5, 0.5, 0.5, 1.0), stroke=(0.0, 0.0, 0.0, 1.0), n_threads=1, model="cairo",
                  rotated_walls="cairo"):
    """Restated ``World.generate_image`` -> (n_worlds, H, W, 4) uint8.

    model "cairo": cairo_oracle.c, the published cairo pipeline (arc -> Beziers -> flattening -> tor scan
    converter -> span compositing); model "area": render_oracle.c, exact shape-pixel areas (kept to measure the
    difference between the two).  rotated_walls "area" (with model "cairo"): non-rectilinear wall quads are drawn as
    the engine draws them -- solid colour through the exact pixel-quad area -- everything else the cairo way.
    wall (n_worlds, n_walls, 4, 2); pos (n_worlds, n_balls, 2); rad (n_worlds, n_balls);
    wall_rgba (n_walls, 4) default red; fill / stroke / s_thick per ball slot or shared.
    """
    L = lib()
    fn = {"cairo": L.cairo_render_frames, "area": L.oracle_render_frames}[model]
    ctypes.c_int.in_dll(L, "cairo_rotated_wall_model").value = {"cairo": 0, "area": 1}[rotated_walls]
    wall = np.ascontiguousarray(wall, np.float64)
    pos = np.ascontiguousarray(pos, np.float64)
    nw, nb = pos.shape[:2]
    nwl = wall.shape[1]
    rad = np.ascontiguousarray(np.broadcast_to(rad, (nw, nb)), np.float64)
    W, H = int(canvas[0]), int(canvas[1])
    order = np.ascontiguousarray(np.arange(nwl + nb) if order is None else order, np.int32)
    wall_kind = np.ascontiguousarray(np.zeros(nwl) if wall_kind is None else wall_kind, np.int32)
    if wall_rgba is None or nwl == 0:
        wall_rgba = np.array([1.0, 0.0, 0.0, 1.0])
    wall_rgba = np.ascontiguousarray(np.broadcast_to(np.asarray(wall_rgba, np.float32), (max(nwl, 1), 4)), np.float32)
    st = np.ascontiguousarray(np.broadcast_to(s_thick, (nb,)), np.int32)
    fl = np.ascontiguousarray(np.broadcast_to(np.asarray(fill, np.float32), (nb, 4)), np.float32)
    sk = np.ascontiguousarray(np.broadcast_to(np.asarray(stroke, np.float32), (nb, 4)), np.float32)
    out = np.empty((nw, H, W, 4), np.uint8)
    fp = ctypes.POINTER(ctypes.c_float)
    fn(ctypes.c_int(nw), ctypes.c_int(nb), ctypes.c_int(nwl), _ip(order), _dp(wall),
       _ip(wall_kind), wall_rgba.ctypes.data_as(fp), _dp(pos), _dp(rad), _ip(st),
       fl.ctypes.data_as(fp), sk.ctypes.data_as(fp), ctypes.c_int(W), ctypes.c_int(H),
       out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.c_int(n_threads))
    ctypes.c_int.in_dll(L, "cairo_rotated_wall_model").value = 0
    return out


def ball_sprite(r, s_thick=2, fill=(0.5, 0.5, 0.5, 1.0), stroke=(0.0, 0.0, 0.0, 1.0), model="cairo"):
    """get_ball_im restated: (sz, sz, 4) premultiplied uint8."""
    L = lib()
    fn = {"cairo": L.cairo_ball_sprite, "area": L.oracle_ball_sprite}[model]
    sz = 2 * (int(r) + int(s_thick))
    out = np.zeros((sz, sz, 4), np.uint8)
    fl = np.asarray(fill, np.float32)
    sk = np.asarray(stroke, np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    fn(ctypes.c_int(int(r)), ctypes.c_int(int(s_thick)), fl.ctypes.data_as(fp),
       sk.ctypes.data_as(fp), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


_u8p = ctypes.POINTER(ctypes.c_uint8)


def cairo_disc_mask(cx, cy, r, w, h):
    """a8 coverage of ``arc(cx, cy, r, 0, 2 pi); fill`` on a (h, w) surface (cairo_oracle.c)."""
    out = np.zeros((h, w), np.uint8)
    lib().cairo_disc_mask(ctypes.c_double(cx), ctypes.c_double(cy), ctypes.c_double(r), ctypes.c_int(w),
                          ctypes.c_int(h), out.ctypes.data_as(_u8p))
    return out


def cairo_ring_mask(cx, cy, r, line_width, w, h):
    """a8 coverage of ``arc(...); set_line_width(line_width); stroke``; also returns the cusp count (0)."""
    out = np.zeros((h, w), np.uint8)
    c = lib().cairo_ring_mask(ctypes.c_double(cx), ctypes.c_double(cy), ctypes.c_double(r),
                              ctypes.c_double(line_width), ctypes.c_int(w), ctypes.c_int(h), out.ctypes.data_as(_u8p))
    return out, int(c)


def cairo_quad_mask(quad, w, h):
    """a8 coverage of the filled quad (4 x (x, y)) on a (h, w) surface."""
    q = np.ascontiguousarray(quad, np.float64)
    out = np.zeros((h, w), np.uint8)
    lib().cairo_quad_mask(_dp(q), ctypes.c_int(w), ctypes.c_int(h), out.ctypes.data_as(_u8p))
    return out


def cairo_circle_flatten(cx, cy, r):
    """Vertices (24.8 fixed point, as integers) of the polygon cairo fills for a full-circle arc."""
    xy = np.zeros((4096, 2), np.int32)
    n = lib().cairo_circle_flatten(ctypes.c_double(cx), ctypes.c_double(cy), ctypes.c_double(r), ctypes.c_int(4096),
                                   _ip(xy))
    return xy[:n]


def cairo_block_sprite(verts, rgba=(1.0, 0.0, 0.0, 1.0)):
    """get_block_im for the wall with world-space vertices ``verts`` (4, 2): (h, w, 4) premultiplied uint8."""
    v = np.ascontiguousarray(verts, np.float64)
    col = np.asarray(rgba, np.float32)
    wh = np.zeros(2, np.int32)
    buf = np.zeros((2048, 2048, 4), np.uint8)
    lib().cairo_block_sprite(_dp(v), col.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.c_int(2048),
                             ctypes.c_int(2048), _ip(wh), buf.ctypes.data_as(_u8p))
    w, h = int(wh[0]), int(wh[1])
    return buf.reshape(-1)[: w * h * 4].reshape(h, w, 4).copy()


def disc_box_area(R, x0, x1, y0, y1):
    return lib().oracle_disc_box_area(ctypes.c_double(R), ctypes.c_double(x0), ctypes.c_double(x1),
                                      ctypes.c_double(y0), ctypes.c_double(y1))


def quad_pixel_area(quad, px, py):
    q = np.ascontiguousarray(quad, np.float64)
    return lib().oracle_quad_pixel_area(_dp(q), ctypes.c_double(px), ctypes.c_double(py))


def trig(which: str, x: float, mode: int) -> float:
    return lib().oracle_trig({"sin": 0, "asin": 1, "acos": 2}[which], float(x), int(mode))
