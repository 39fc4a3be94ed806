"""WorldBatch -- N independent billiards worlds resident in HBM.

Thin Python view of the C ABI (include/pyphy_b200.h).  This is the batched
counterpart of ``pm.World`` + ``pm.Dynamics`` (reference primitives.py:535-805,
856-1008): the façade classes in ``pyphy_engine_b200.primitives`` are the
``n_worlds == 1`` case of this class, ``pyphy_engine_b200.dataio`` uses it with
one world per sequence.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _capi
from ._capi import PPB_F32, PPB_F64, BallStyle, Config, Event, WallStyle, check

def premul_rgba8(rgba):
    """Colour -> premultiplied 8-bit (r, g, b, a) the way cairo/pixman quantise: 16-bit then >> 8
    (same arithmetic as h_premul_rgba8 in csrc/ppb_api.cu)."""
    def q(v):
        v = min(1.0, max(0.0, float(v)))
        return int(v * 65535.0 + 0.5) >> 8
    a = float(np.float32(rgba[3]))
    return tuple(q(float(np.float32(c)) * a) for c in rgba[:3]) + (q(a),)


_DTYPES = {"f32": PPB_F32, "float32": PPB_F32, "fp32": PPB_F32, "f64": PPB_F64, "float64": PPB_F64, "fp64": PPB_F64}


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _stream_handle(stream, device=None) -> int:
    """cudaStream_t of a torch stream / raw handle; None = torch's current stream (every WorldBatch
    method defaults to it, so state transfers, steps and read-backs of one thread are stream-ordered)."""
    if stream is None:
        import torch
        return int(torch.cuda.current_stream(device).cuda_stream)
    if hasattr(stream, "cuda_stream"):
        return int(stream.cuda_stream)
    return int(stream)


def collision_flags(traj, thresh: float = 0.03, stream=None):
    """utils.detect_collision's test on a float32 CUDA trajectory (n_steps, n_items, 2):
    (n_items, n_steps - 2) uint8, 1 where |diff(diff(pos))|^2 >= thresh * |diff(pos)|^2 (utils.py:38-49)."""
    import torch
    assert traj.is_cuda and traj.is_contiguous() and traj.dtype == torch.float32 and traj.shape[-1] == 2
    n_steps = int(traj.shape[0])
    n_items = int(traj.numel() // (2 * n_steps))
    flags = torch.empty((n_items, n_steps - 2), dtype=torch.uint8, device=traj.device)
    with torch.cuda.device(traj.device):
        check(_capi.load().ppb_collision_flags_async(ctypes.c_void_p(traj.data_ptr()), n_items, n_steps, float(thresh),
                                                     ctypes.c_void_p(flags.data_ptr()), _stream_handle(stream, traj.device)))
    return flags


class WorldBatch:
    """``n_worlds`` worlds, each with ``n_walls`` wall slots and ``n_balls`` ball slots.

    dtype "f64": the reference's fp64 operation order (bit-exact collision events);
    dtype "f32": throughput mode.  ``order`` is the object insertion order shared
    by all worlds (wall q -> q, ball b -> n_walls + b), default walls then balls.
    """

    def __init__(self, n_worlds: int, n_balls: int, n_walls: int, dtype: str = "f32",
                 canvas: Sequence[int] = (0, 0), dt: float = 0.01, g: Sequence[float] = (0.0, 0.0),
                 order: Optional[Sequence[int]] = None, device: int = 0, event_capacity: int = 0,
                 friction: float = 0.0, max_substeps: int = 0):
        self._lib = _capi.load()
        self._h = ctypes.c_void_p()
        self.n_worlds, self.n_balls, self.n_walls = int(n_worlds), int(n_balls), int(n_walls)
        self.dtype = dtype
        self.canvas_w, self.canvas_h = int(canvas[0]), int(canvas[1])
        self.device = int(device)
        self.dt = float(dt)
        cfg = Config()
        cfg.n_worlds, cfg.n_balls, cfg.n_walls = self.n_worlds, self.n_balls, self.n_walls
        cfg.dtype = _DTYPES[dtype]
        cfg.canvas_w, cfg.canvas_h = self.canvas_w, self.canvas_h
        cfg.device = self.device
        cfg.dt, cfg.gx, cfg.gy, cfg.friction = float(dt), float(g[0]), float(g[1]), float(friction)
        cfg.event_capacity = int(event_capacity)
        self.event_capacity = int(event_capacity)
        cfg.max_substeps = int(max_substeps)
        self._order = None
        if order is not None:
            self._order = np.ascontiguousarray(order, np.int32)
            if self._order.shape != (self.n_walls + self.n_balls,):
                raise ValueError("order must have n_walls + n_balls entries")
            cfg.order = _ip(self._order)
        check(self._lib.ppb_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.steps_done = 0
        self._np_dtype = np.float64 if cfg.dtype == PPB_F64 else np.float32

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ppb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state in
    def set_walls(self, wall, world0: int = 0, stream=None):
        """wall: (count, n_walls, 4, 2) vertices (Bbox order, geometry.py:471-505)."""
        wall = np.ascontiguousarray(wall, np.float64)
        count = wall.shape[0] if self.n_walls else self.n_worlds - world0
        if self.n_walls and wall.shape[1:] != (self.n_walls, 4, 2):
            raise ValueError("wall must be (count, %d, 4, 2)" % self.n_walls)
        check(self._lib.ppb_set_walls(self._h, world0, count, _dp(wall), _stream_handle(stream, self.device)))

    def set_balls(self, pos, vel, rad, mass, world0: int = 0, stream=None):
        """pos, vel: (count, n_balls, 2); rad, mass: (count, n_balls) or broadcastable.
        Resets the collision caches of these worlds (Dynamics._include_object)."""
        pos = np.ascontiguousarray(pos, np.float64)
        count = pos.shape[0]
        shp = (count, self.n_balls)
        if pos.shape != shp + (2,):
            raise ValueError("pos must be (count, %d, 2)" % self.n_balls)
        vel = np.ascontiguousarray(np.broadcast_to(np.asarray(vel, np.float64), shp + (2,)))
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(rad, np.float64), shp))
        mass = np.ascontiguousarray(np.broadcast_to(np.asarray(mass, np.float64), shp))
        check(self._lib.ppb_set_balls(self._h, world0, count, _dp(pos), _dp(vel), _dp(rad), _dp(mass), _stream_handle(stream, self.device)))

    def sample_rect_worlds(self, seed: int = 0, world0: int = 0, count: Optional[int] = None, world_index0: int = 0,
                           arena=700, w_thick=30, mn_wlen=300, mx_wlen=550, mn_ball=25, mx_ball=25, mn_force=3e4,
                           mx_force=8e4, force_t=1.0, max_tries=500, w_theta=None, stream=None) -> int:
        """Fill worlds [world0, world0+count) with random tables drawn ON THE DEVICE from the DataSaver
        distribution (dataio.py:198-416): rectangular (isRect=True) by default, rhombus tables with the
        side angle drawn from ``w_theta`` (degrees, up to 4 values; isRect=False, wTheta) otherwise.
        Returns the number of worlds that could not be filled.  World w depends on (seed, world_index0 + w) only."""
        count = self.n_worlds - world0 if count is None else count
        thetas = [] if w_theta is None else ([float(w_theta)] if np.isscalar(w_theta) else [float(t) for t in w_theta])
        if len(thetas) > 4:
            raise ValueError("at most 4 rhombus angles")
        sp = _capi.SamplerParams(float(arena), float(w_thick), float(mn_wlen), float(mx_wlen), float(mn_ball),
                                 float(mx_ball), float(mn_force), float(mx_force), float(force_t), int(seed),
                                 int(world_index0), int(max_tries), len(thetas),
                                 (ctypes.c_double * 4)(*(thetas + [0.0] * (4 - len(thetas)))))
        nf = ctypes.c_int32(0)
        check(self._lib.ppb_sample_rect_worlds(self._h, world0, count, ctypes.byref(sp), ctypes.byref(nf), _stream_handle(stream, self.device)))
        return nf.value

    def get_walls(self, world0: int = 0, count: Optional[int] = None, stream=None):
        """(count, n_walls, 4, 2) wall vertices read back from the device."""
        count = self.n_worlds - world0 if count is None else count
        wall = np.empty((count, self.n_walls, 4, 2), np.float64)
        check(self._lib.ppb_get_walls(self._h, world0, count, _dp(wall), _stream_handle(stream, self.device)))
        return wall

    def get_radius_mass(self, world0: int = 0, count: Optional[int] = None, stream=None):
        """(radius, mass) per ball slot, (count, n_balls) each; radius < 0 marks an empty slot."""
        count = self.n_worlds - world0 if count is None else count
        rad = np.empty((count, self.n_balls), np.float64)
        mass = np.empty((count, self.n_balls), np.float64)
        check(self._lib.ppb_get_radius_mass(self._h, world0, count, _dp(rad), _dp(mass), _stream_handle(stream, self.device)))
        return rad, mass

    def update_balls(self, pos=None, vel=None, world0: int = 0, stream=None):
        """ball.set_position / set_velocity: caches are left alone (primitives.py:462-466)."""
        ref = pos if pos is not None else vel
        if ref is None:
            return
        count = np.asarray(ref).shape[0]
        p = None if pos is None else np.ascontiguousarray(pos, np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, np.float64)
        for a in (p, v):
            if a is not None and a.shape != (count, self.n_balls, 2):
                raise ValueError("expected (count, n_balls, 2)")
        check(self._lib.ppb_update_balls(self._h, world0, count, _dp(p), _dp(v), _stream_handle(stream, self.device)))

    def apply_force(self, force, force_t: float, world0: int = 0, stream=None):
        """Dynamics.apply_force for every ball of ``count`` worlds (primitives.py:580-594): ``force`` is
        (count, n_balls, 2); vel += 0.5 * force_t**2 * force / mass, caches untouched (quirk Q7)."""
        f = np.ascontiguousarray(force, np.float64)
        if f.ndim != 3 or f.shape[1:] != (self.n_balls, 2):
            raise ValueError("expected (count, n_balls, 2)")
        check(self._lib.ppb_apply_force(self._h, world0, f.shape[0], _dp(f), float(force_t),
                                        _stream_handle(stream, self.device)))

    def requeue(self, world0: int = 0, count: Optional[int] = None, stream=None):
        """Dynamics.add_all_dynamic_collision_queue (primitives.py:697-700)."""
        count = self.n_worlds - world0 if count is None else count
        check(self._lib.ppb_requeue(self._h, world0, count, _stream_handle(stream, self.device)))

    def set_styles(self, ball_styles, wall_styles, r_min: int, r_max: int, stream=None):
        """ball_styles: per slot (s_thick, fill rgba, stroke rgba); wall_styles: per slot (kind, fill rgba)."""
        bs = (BallStyle * self.n_balls)()
        for i, (st, fill, stroke) in enumerate(ball_styles):
            bs[i].s_thick = int(st)
            bs[i].fill[:] = [float(c) for c in fill]
            bs[i].stroke[:] = [float(c) for c in stroke]
        ws = (WallStyle * max(1, self.n_walls))()
        for i, (kind, fill) in enumerate(wall_styles):
            ws[i].kind = int(kind)
            ws[i].fill[:] = [float(c) for c in fill]
        check(self._lib.ppb_set_styles(self._h, bs, ws, int(r_min), int(r_max), _stream_handle(stream, self.device)))

    def set_gravity(self, gx: float, gy: float):
        """Dynamics.set_g (primitives.py:566); collision caches are left alone."""
        check(self._lib.ppb_set_gravity(self._h, float(gx), float(gy)))

    def read_sprite(self, slot: int, radius: int, stream=None):
        """get_ball_im (primitives.py:33-61) as built on the device: (sz, sz, 4) uint8, premultiplied."""
        sz = ctypes.c_int32(0)
        check(self._lib.ppb_read_sprite(self._h, int(slot), int(radius), None, 0, ctypes.byref(sz), _stream_handle(stream, self.device)))
        out = np.empty((sz.value, sz.value, 4), np.uint8)
        check(self._lib.ppb_read_sprite(self._h, int(slot), int(radius), ctypes.c_void_p(out.ctypes.data), out.nbytes,
                                        ctypes.byref(sz), _stream_handle(stream, self.device)))
        return out

    # ------------------------------------------------------------------ state out
    def get_balls(self, world0: int = 0, count: Optional[int] = None, stream=None):
        count = self.n_worlds - world0 if count is None else count
        pos = np.empty((count, self.n_balls, 2), np.float64)
        vel = np.empty((count, self.n_balls, 2), np.float64)
        check(self._lib.ppb_get_balls(self._h, world0, count, _dp(pos), _dp(vel), _stream_handle(stream, self.device)))
        return pos, vel

    def get_collision_cache(self, world0: int = 0, count: Optional[int] = None, stream=None):
        count = self.n_worlds - world0 if count is None else count
        tcol = np.empty((count, self.n_balls), np.float64)
        partner = np.empty((count, self.n_balls), np.int32)
        check(self._lib.ppb_get_collision_cache(self._h, world0, count, _dp(tcol), _ip(partner), _stream_handle(stream, self.device)))
        return tcol, partner

    def scan(self, stream=None):
        """Drain the collision queue without stepping (time_to_collide_all for every queued ball)."""
        check(self._lib.ppb_scan_async(self._h, _stream_handle(stream, self.device)))

    def get_collision_response(self, world0: int = 0, count: Optional[int] = None, stream=None):
        """(nrmlCol_, futureVel_) per ball: (count, n_balls, 2) each."""
        count = self.n_worlds - world0 if count is None else count
        nrml = np.empty((count, self.n_balls, 2), np.float64)
        fv = np.empty((count, self.n_balls, 2), np.float64)
        check(self._lib.ppb_get_collision_response(self._h, world0, count, _dp(nrml), _dp(fv), _stream_handle(stream, self.device)))
        return nrml, fv

    def get_status(self, world0: int = 0, count: Optional[int] = None, stream=None):
        """(status, steps_done) per world; for a halted world steps_done is the failing step."""
        count = self.n_worlds - world0 if count is None else count
        status = np.empty((count,), np.int32)
        steps = np.empty((count,), np.int32)
        check(self._lib.ppb_get_status(self._h, world0, count, _ip(status), _ip(steps), _stream_handle(stream, self.device)))
        return status, steps

    def read_events(self, stream=None):
        """(k, 4) int32 rows (world, step, ball, partner) sorted by (world, call order)."""
        n = ctypes.c_int64(0)
        check(self._lib.ppb_read_events(self._h, None, 0, ctypes.byref(n), _stream_handle(stream, self.device)))
        total = n.value
        m = min(total, self.event_capacity)          # the ring keeps the first event_capacity records
        buf = (Event * max(1, m))()
        check(self._lib.ppb_read_events(self._h, buf, m, ctypes.byref(n), _stream_handle(stream, self.device)))
        raw = np.frombuffer(buf, dtype=np.dtype([("world", "<i4"), ("seq", "<i4"), ("step", "<i4"),
                                                  ("ball", "<i2"), ("partner", "<i2")]), count=m)
        idx = np.lexsort((raw["seq"], raw["world"]))
        raw = raw[idx]
        out = np.stack([raw["world"], raw["step"], raw["ball"].astype(np.int32), raw["partner"].astype(np.int32)], 1)
        return out.astype(np.int32), total

    def clear_events(self, stream=None):
        check(self._lib.ppb_clear_events(self._h, _stream_handle(stream, self.device)))

    def counters(self, stream=None):
        """dict(launches, collide_world_steps, collide_iterations, events, rescans, pair_tests)."""
        out = (ctypes.c_int64 * 8)()
        check(self._lib.ppb_get_counters(self._h, out, _stream_handle(stream, self.device)))
        per_rescan = self.n_balls * (4 * self.n_walls + self.n_balls - 1)
        return dict(launches=out[0], collide_world_steps=out[1], collide_iterations=out[2], events=out[3],
                    rescans=out[4], pair_tests=out[4] * per_rescan, d2h_packed_bytes=out[5], d2h_raw_bytes=out[6],
                    unpack_us=out[7])

    # ------------------------------------------------------------------ stepping
    def step_async(self, n_steps: int = 1, traj=None, frames=None, render_every: int = 1, stream=None):
        """n_steps x Dynamics.step() on `stream` (default: torch's current stream).

        traj: None or a CUDA tensor (n_steps, n_worlds, n_balls, 2) of the batch dtype;
        frames: None or a CUDA uint8 tensor (n_steps // render_every, n_worlds, H, W, 4).
        """
        tp = ctypes.c_void_p(traj.data_ptr()) if traj is not None else None
        fp = ctypes.c_void_p(frames.data_ptr()) if frames is not None else None
        if traj is not None:
            assert traj.is_cuda and traj.is_contiguous() and tuple(traj.shape) == (n_steps, self.n_worlds, self.n_balls, 2)
        if frames is not None:
            assert frames.is_cuda and frames.is_contiguous()
            assert tuple(frames.shape) == (n_steps // render_every, self.n_worlds, self.canvas_h, self.canvas_w, 4)
        check(self._lib.ppb_step_async(self._h, int(n_steps), tp, fp, int(render_every), _stream_handle(stream, self.device)))
        self.steps_done += int(n_steps)

    def step_ring_async(self, n_steps: int, frames, render_every: int = 1, traj=None, stream=None):
        """n_steps x (step + render) with the frames written round-robin into the slots of
        ``frames`` (ring_slots, n_worlds, H, W, 4): bounded frame memory for long runs."""
        assert frames.is_cuda and frames.is_contiguous() and frames.dtype.itemsize == 1
        assert tuple(frames.shape[1:]) == (self.n_worlds, self.canvas_h, self.canvas_w, 4)
        tp = ctypes.c_void_p(traj.data_ptr()) if traj is not None else None
        check(self._lib.ppb_step_ring_async(self._h, int(n_steps), tp, ctypes.c_void_p(frames.data_ptr()),
                                            int(frames.shape[0]), int(render_every), _stream_handle(stream, self.device)))
        self.steps_done += int(n_steps)

    def step(self, n_steps: int = 1, record_traj: bool = False, render_every: int = 0, stream=None):
        """Synchronous convenience: returns (traj, frames) as CUDA tensors (or None)."""
        import torch
        dev = torch.device("cuda", self.device)
        traj = frames = None
        if record_traj:
            traj = torch.empty((n_steps, self.n_worlds, self.n_balls, 2), device=dev,
                               dtype=torch.float64 if self._np_dtype is np.float64 else torch.float32)
        if render_every:
            frames = torch.empty((n_steps // render_every, self.n_worlds, self.canvas_h, self.canvas_w, 4),
                                 device=dev, dtype=torch.uint8)
        self.step_async(n_steps, traj, frames, render_every or 1, stream)
        torch.cuda.synchronize(dev)
        return traj, frames

    def render(self, out=None, stream=None):
        """World.generate_image() for every world -> CUDA uint8 tensor (n_worlds, H, W, 4)."""
        import torch
        if out is None:
            out = torch.empty((self.n_worlds, self.canvas_h, self.canvas_w, 4), device=torch.device("cuda", self.device),
                              dtype=torch.uint8)
        check(self._lib.ppb_render_async(self._h, ctypes.c_void_p(out.data_ptr()), _stream_handle(stream, self.device)))
        return out

    def crop(self, frames, traj, crop_sz: int, mode: int = 0, stream=None):
        """Ball-centred white-padded crops (dataio.py:160-182) of frames (F, n_worlds, H, W, 4) around the centres
        traj (F, n_worlds, n_balls, 2): returns (crops (F, n_worlds, n_balls, c, c, 3) uint8, positions in crop
        coordinates (F, n_worlds, n_balls, 2) float32), both CUDA tensors.  mode 1 = CustomWorld.get_glimpse
        rounding (dataio.py:505-518; positions returned unchanged)."""
        import torch
        F = int(frames.shape[0])
        assert frames.is_cuda and frames.is_contiguous() and tuple(frames.shape[1:]) == (self.n_worlds, self.canvas_h, self.canvas_w, 4)
        assert traj.is_cuda and traj.is_contiguous() and tuple(traj.shape) == (F, self.n_worlds, self.n_balls, 2)
        dev = frames.device
        crops = torch.empty((F, self.n_worlds, self.n_balls, crop_sz, crop_sz, 3), dtype=torch.uint8, device=dev)
        pos = torch.empty((F, self.n_worlds, self.n_balls, 2), dtype=torch.float32, device=dev)
        check(self._lib.ppb_crop_async(self._h, ctypes.c_void_p(frames.data_ptr()), ctypes.c_void_p(traj.data_ptr()), F,
                                       int(crop_sz), int(mode), ctypes.c_void_p(crops.data_ptr()), ctypes.c_void_p(pos.data_ptr()),
                                       _stream_handle(stream, self.device)))
        return crops, pos

    def resize_crops(self, crops, proc_sz: int, pos=None, traj=None, want_vel: bool = False, stream=None):
        """The procSz step of DataSaver.fetch (dataio.py:183-189) on the device: PIL-bilinear resize (what
        scipy.misc.imresize does) of crops (F, n_worlds, n_balls, c, c, 3) -> (F, n_worlds, n_balls, p, p, 3); `pos`
        (crop coordinates from ``crop``) is multiplied in place by p / c; with want_vel the displacements between
        consecutive frames of `traj`, times p / c, are returned as (F, n_worlds, n_balls, 2) float32 (row i-1 holds
        frame i minus frame i-1, the last row is zero).  Returns (resized, pos, vel or None)."""
        import torch
        F, c = int(crops.shape[0]), int(crops.shape[3])
        assert crops.is_cuda and crops.is_contiguous() and tuple(crops.shape) == (F, self.n_worlds, self.n_balls, c, c, 3)
        out = torch.empty((F, self.n_worlds, self.n_balls, proc_sz, proc_sz, 3), dtype=torch.uint8, device=crops.device)
        vel = None
        if want_vel:
            assert traj is not None and traj.is_cuda and traj.is_contiguous() and tuple(traj.shape) == (F, self.n_worlds, self.n_balls, 2)
            vel = torch.zeros((F, self.n_worlds, self.n_balls, 2), dtype=torch.float32, device=crops.device)
        if pos is not None:
            assert pos.is_cuda and pos.is_contiguous() and pos.dtype == torch.float32
        check(self._lib.ppb_resize_crops_async(
            self._h, ctypes.c_void_p(crops.data_ptr()), ctypes.c_void_p(traj.data_ptr()) if traj is not None else None, F, c,
            int(proc_sz), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(pos.data_ptr()) if pos is not None else None,
            ctypes.c_void_p(vel.data_ptr()) if vel is not None else None, _stream_handle(stream, self.device)))
        return out, pos, vel

    def horizon(self, traj, lookahead: int, stream=None):
        """DynamicsHorizon.get_data's look-ahead matrix (primitives.py:839-852) for every row of a recorded
        trajectory (R, n_worlds, n_balls, 2) of the batch dtype (row 0 = positions before the first step):
        (R - lookahead, n_worlds, n_balls, 2 * lookahead) float32, [t, w, b, 2k + c] = traj[t+k+1] - traj[t+k]."""
        import torch
        R = int(traj.shape[0])
        assert traj.is_cuda and traj.is_contiguous() and tuple(traj.shape) == (R, self.n_worlds, self.n_balls, 2)
        assert traj.dtype == (torch.float64 if self._np_dtype is np.float64 else torch.float32)
        out = torch.empty((R - lookahead, self.n_worlds, self.n_balls, 2 * lookahead), dtype=torch.float32, device=traj.device)
        check(self._lib.ppb_horizon_async(self._h, ctypes.c_void_p(traj.data_ptr()), R, int(lookahead),
                                          ctypes.c_void_p(out.data_ptr()), _stream_handle(stream, self.device)))
        return out

    def simulate_host(self, n_steps: int, traj_host=None, frames_host=None, render_every: int = 1, stream=None,
                      wait: bool = True):
        """Host-buffer path (DataSaver.fetch analogue): copies are inside the call.

        traj_host: None or float32 array (n_steps, n_worlds, n_balls, 2);
        frames_host: None or uint8 array (n_steps // render_every, n_worlds, H, W, 4).
        numpy arrays or pinned CPU torch tensors are accepted.
        wait=False returns once everything is enqueued (ppb_simulate_host_async); the host buffers are complete
        after ``host_sync()`` and must not be reused by another call before that.
        """
        def _ptr(a):
            if a is None:
                return None
            if hasattr(a, "data_ptr"):
                return ctypes.c_void_p(a.data_ptr())
            return ctypes.c_void_p(a.ctypes.data)
        fn = self._lib.ppb_simulate_host if wait else self._lib.ppb_simulate_host_async
        check(fn(self._h, int(n_steps), _ptr(traj_host), _ptr(frames_host), int(render_every),
                 _stream_handle(stream, self.device)))
        self.steps_done += int(n_steps)

    def download_frames(self, frames, out=None, stream=None):
        """Device frames (..., H, W, 4) uint8 -> host array of the same shape through the packed download
        (ppb_download_frames): run-length tokens over PCIe, pixels rebuilt by host threads.  `out`: a numpy array or a
        pinned CPU torch tensor to fill (default: a new numpy array)."""
        assert frames.is_cuda and frames.is_contiguous() and frames.dtype.itemsize == 1
        assert tuple(frames.shape[-3:]) == (self.canvas_h, self.canvas_w, 4)
        n = int(np.prod(frames.shape[:-3])) if frames.dim() > 3 else 1
        if out is None:
            out = np.empty(tuple(frames.shape), np.uint8)
        ptr = out.data_ptr() if hasattr(out, "data_ptr") else out.ctypes.data
        check(self._lib.ppb_download_frames(self._h, ctypes.c_void_p(frames.data_ptr()), n, ctypes.c_void_p(ptr),
                                            _stream_handle(stream, self.device)))
        return out

    def host_sync(self, keep_in_flight: int = 0, stream=None):
        """Wait for the steps and device -> host copies of earlier ``simulate_host(..., wait=False)`` calls:
        all of them (0) or all but the most recent one (1)."""
        check(self._lib.ppb_host_sync(self._h, int(keep_in_flight), _stream_handle(stream, self.device)))

    # ------------------------------------------------------------------ profiling
    def set_profiling(self, on):
        """False/0 off, True/1 every kernel, 2 the render kernel only, 3 the advance (physics) kernel only."""
        check(self._lib.ppb_set_profiling(self._h, int(on)))

    def kernel_times(self):
        """(ms[3], launches[3]) of the last stepping call: [0] advance (physics), [1] unused, [2] render."""
        ms = (ctypes.c_float * 3)()
        n = (ctypes.c_int32 * 3)()
        check(self._lib.ppb_get_kernel_times(self._h, ms, n))
        return list(ms), list(n)

    def device_ptr(self, which: int):
        p = ctypes.c_void_p()
        n = ctypes.c_int64()
        check(self._lib.ppb_device_ptr(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value
