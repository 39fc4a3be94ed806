"""GPU parity tests (frames): render kernel through the C ABI vs. oracle/cairo_oracle.c, the restatement of the
published cairo pipeline the reference draws through (arc -> Beziers -> flattening -> stroker -> tor scan
converter -> span compositing; see that file's header).

Frame parity against libcairo's own output is UNPINNED (no pycairo/libcairo in the reference tree or in this
image, no golden images in the reference).  What is tested here:
  * ball sprites, axis-aligned walls (every DataSaver rect table, BASELINE configs 1-4), clipping, insertion
    order, translucent colours: identical bytes;
  * rotated wall quads (rhombus / n-gon tables): the engine paints the wall colour through the exact
    pixel-quad area instead of cairo's sampled coverage of the quad and of its pre-rendered stroked sprite.
    Against that stated model |diff| <= RENDER_TOL (fp32 on the device, fp64 in the oracle); against the cairo
    restatement |diff| <= CAIRO_WALL_TOL, only on the walls' edge pixels (measured: tests/test_cairo_oracle.py).
"""
import numpy as np
import pytest

import oracle
from conftest import load_golden

pytestmark = pytest.mark.gpu

RENDER_TOL = 1  # LSB per channel, rotated-quad edge pixels only
CAIRO_WALL_TOL = 32  # per channel, of 255; measured maximum over the golden rotated tables: 30

GRAY = (0.5, 0.5, 0.5, 1.0)
BLACK = (0.0, 0.0, 0.0, 1.0)
RED = (1.0, 0.0, 0.0, 1.0)


def _render_both(wall, pos, rad, canvas, s_thick, wall_kind=None, wall_rgba=None, fill=GRAY, stroke=BLACK,
                 order=None, dtype="f32", rotated_walls="cairo"):
    from pyphy_engine_b200.engine import WorldBatch
    wall = np.asarray(wall, np.float64)
    pos = np.asarray(pos, np.float64)
    nw, nb = pos.shape[:2]
    nwl = wall.shape[1]
    rad = np.broadcast_to(np.asarray(rad, np.float64), (nw, nb))
    wb = WorldBatch(nw, nb, nwl, dtype=dtype, canvas=canvas, order=order)
    wb.set_walls(wall)
    wb.set_balls(pos, np.zeros_like(pos), rad, np.ones((nw, nb)))
    kinds = [0] * nwl if wall_kind is None else list(wall_kind)
    cols = [RED] * nwl if wall_rgba is None else list(wall_rgba)
    live = rad[rad >= 0]
    wb.set_styles([(s_thick, fill, stroke)] * nb, list(zip(kinds, cols)), int(live.min()), int(live.max()))
    dev = wb.render().cpu().numpy()
    ref = oracle.render_frames(wall, pos, rad, canvas, order=order, wall_kind=kinds, wall_rgba=np.array(cols),
                               s_thick=s_thick, fill=fill, stroke=stroke, n_threads=8, rotated_walls=rotated_walls)
    return dev, ref


def test_frames_64px_8_balls_identical():
    """BASELINE config 4 geometry: 64x64 canvas, 8 balls of radius 3, axis-aligned cage."""
    from pyphy_engine_b200 import sampler
    W = sampler.sample_rect_worlds(1024, 8, seed=2, **sampler.scaled64_params(8))
    dev, ref = _render_both(W["wall"], W["pos"], W["rad"], (64, 64), 1)
    assert dev.shape == ref.shape == (1024, 64, 64, 4)
    assert np.array_equal(dev, ref)
    assert (ref != 255).any()


def test_frames_after_stepping_identical():
    """Frames produced inside the step loop (fp32 physics) equal the oracle's rendering of the
    positions the engine reports for the same step."""
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    W = sampler.sample_rect_worlds(256, 4, seed=3, **sampler.scaled64_params(4))
    wb = WorldBatch(256, 4, 4, dtype="f32", canvas=(64, 64))
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"] * 20.0, W["rad"], W["mass"])   # faster balls: visible motion
    wb.set_styles([(1, GRAY, BLACK)] * 4, [(0, RED)] * 4, 3, 3)
    traj, frames = wb.step(40, record_traj=True, render_every=10)
    traj = traj.cpu().numpy().astype(np.float64)
    frames = frames.cpu().numpy()
    assert frames.shape == (4, 256, 64, 64, 4)
    for k in range(4):
        ref = oracle.render_frames(W["wall"], traj[10 * k + 9], W["rad"], (64, 64), s_thick=1, n_threads=8)
        assert np.array_equal(frames[k], ref)
    assert not np.array_equal(frames[0], frames[3])


def test_frames_native_700_identical():
    """ICLR16 native size: arena 700, r = 25, sThick = 2 (dataio.py:336, 435-446)."""
    from pyphy_engine_b200 import sampler
    W = sampler.sample_rect_worlds(6, 4, seed=5)
    dev, ref = _render_both(W["wall"], W["pos"], W["rad"], (700, 700), 2)
    assert np.array_equal(dev, ref)


def test_frames_odd_canvas_and_clipping():
    """arenaSz = 667 (dataio.py:26: the canvas width is not a multiple of 4) and balls that stick
    out of the canvas on every side (primitives.py:429-440)."""
    wall = np.array([[[[10, 10], [657, 10], [657, 40], [10, 40]]]], np.float64)
    pos = np.array([[[3.4, 5.5], [665.0, 300.2], [333.5, 664.49], [-30.0, 100.0], [100.5, 100.5]]], np.float64)
    dev, ref = _render_both(wall, pos, 25, (667, 667), 2, dtype="f64")
    assert np.array_equal(dev, ref)


def test_frames_walldef_walls_identical():
    """cairo_trial.create_single_ball_world_gray: WallDef walls (kind 1), gray on white, 640x480."""
    case = [c for c in load_golden("physics_golden.npz") if c.name == "trial_single"][0]
    gray = (0.5, 0.5, 0.5, 1.0)
    dev, ref = _render_both(case["wall"][None], case["pos0"][None], case["rad"][None], (640, 480), 2,
                            wall_kind=[1] * 4, wall_rgba=[gray] * 4, order=case["order"], dtype="f64")
    assert np.array_equal(dev, ref)


@pytest.mark.parametrize("name,canvas", [("diamond", (640, 480)), ("rhomb3_s300", (1600, 1600)),
                                         ("rhomb3_s305", (1600, 1600)), ("ngon3_s800", (800, 800)),
                                         ("ngon6_s801", (800, 800)), ("ngon8_s800", (800, 800))])
def test_frames_rotated_walls(name, canvas):
    case = [c for c in load_golden("physics_golden.npz") if c.name == name][0]
    dev, ref = _render_both(case["wall"][None], case["pos0"][None], case["rad"][None], canvas, 2,
                            order=case["order"], dtype="f64", rotated_walls="area")
    diff = np.abs(dev.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= RENDER_TOL                      # the engine's stated model of a rotated quad
    assert (diff > 0).mean() < 1e-3
    assert (ref[..., 1] == 0).any()     # the red walls are there
    cairo = oracle.render_frames(case["wall"][None], case["pos0"][None], case["rad"][None], canvas, order=case["order"],
                                 s_thick=2, n_threads=8)
    dc = np.abs(dev.astype(np.int32) - cairo.astype(np.int32)).max(axis=-1)
    assert dc.max() <= CAIRO_WALL_TOL
    # pixels that differ are antialiased wall-edge pixels in one rendering or the other, and few
    partly = ((cairo[..., 1] > 0) & (cairo[..., 1] < 255)) | ((dev[..., 1] > 0) & (dev[..., 1] < 255))
    assert not ((dc > 0) & ~partly).any()
    assert (dc > 0).mean() < 0.01


def test_frames_interleaved_order_and_colours():
    """Objects are composited in insertion order (primitives.py:1003-1004): a ball inserted before
    a wall is painted over by it."""
    wall = np.array([[[[20, 20], [60, 20], [60, 60], [20, 60]], [[70, 10], [90, 10], [90, 90], [70, 90]]]], np.float64)
    pos = np.array([[[40.0, 40.0], [80.0, 50.0]]])
    order = [1 + 1, 0, 2 + 1, 1]   # ball-0, wall-0, ball-1, wall-1
    # (a translucent wall is a WallDef wall: GenericWall sprites are always opaque red in the reference, quirk Q19)
    dev, ref = _render_both(wall, pos, 10, (100, 100), 2, order=order, wall_rgba=[(0.2, 0.4, 0.6, 1.0), (0.9, 0.8, 0.1, 0.5)],
                            wall_kind=[0, 1], fill=(0.1, 0.7, 0.3, 0.8), stroke=(0.3, 0.0, 0.9, 1.0))
    assert np.array_equal(dev, ref)
    assert tuple(dev[0, 40, 40, :3]) == tuple(ref[0, 40, 40, :3])
    assert dev[0, 40, 40, 0] == 51        # wall-0 (0.2 -> 51) covers ball-0


def test_sprite_matches_oracle_sprite():
    """get_ball_im restated on the device == restated on the CPU for every radius 1..40."""
    for r, st in ((1, 1), (3, 1), (4, 1), (15, 2), (25, 2), (35, 2), (40, 3)):
        pos = np.array([[[50.0, 50.0]]])
        dev, ref = _render_both(np.zeros((1, 0, 4, 2)), pos, r, (100, 100), st)
        assert np.array_equal(dev, ref), (r, st)
        spr = oracle.ball_sprite(r, st)
        off = r + st
        # OVER a white canvas: out = s + mul(255, 255 - a)
        a = spr[..., 3].astype(np.int32)
        exp = spr[..., 0].astype(np.int32) + (255 - a)
        assert np.array_equal(dev[0, 50 - off:50 + off, 50 - off:50 + off, 0].astype(np.int32), exp)


@pytest.mark.parametrize("n,nb", [(2048, 8), (40000, 3), (20000, 5)])
def test_pipelined_ring_equals_serial_stepping(n, nb):
    """ppb_step_ring_async: render of step k overlaps the physics of step k+1 and the frames go
    round-robin into a ring; the result must equal one-step-at-a-time serial stepping.  The 3- and 5-ball
    batches are large enough for the integrate grid to run in several waves next to the render kernel."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    steps, slots = 9, 3
    W = sampler.sample_rect_worlds(n, nb, seed=12, **sampler.scaled64_params(nb))

    def make():
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
        wb.set_walls(W["wall"])
        wb.set_balls(W["pos"], W["vel"] * 10.0, W["rad"], W["mass"])
        wb.set_styles([(1, GRAY, BLACK)] * nb, [(0, RED)] * 4, 3, 3)
        return wb

    a, b = make(), make()
    ring = torch.zeros((slots, n, 64, 64, 4), dtype=torch.uint8, device="cuda")
    traj = torch.zeros((steps, n, nb, 2), dtype=torch.float32, device="cuda")
    a.step_ring_async(steps, ring, 1, traj=traj)
    torch.cuda.synchronize()
    serial = []
    for s in range(steps):
        t, f = b.step(1, record_traj=True, render_every=1)
        serial.append((t[0].clone(), f[0].clone()))
    for s in range(steps):
        assert torch.equal(traj[s], serial[s][0])
    for s in range(steps - slots, steps):                      # the ring keeps the last `slots` frames
        assert torch.equal(ring[s % slots], serial[s][1]), s
    pa, va = a.get_balls()
    pb, vb = b.get_balls()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb)


@pytest.mark.parametrize("steps,every,slots", [(40, 1, 5), (50, 3, 16), (150, 1, 4), (270, 2, 3), (33, 2, 40)])
def test_chunked_advance_launches_equal_single_steps(steps, every, slots):
    """One advance launch covers up to 64 drawn steps (their position snapshots feed the render launches that follow)
    and the steps after the last drawn one run without snapshots: frames, trajectories and final state must equal
    stepping and rendering one step at a time, whatever the render period and the ring size."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    n, nb = 3000, 8
    W = sampler.sample_rect_worlds(n, nb, seed=21, **sampler.scaled64_params(nb))

    def make():
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
        wb.set_walls(W["wall"])
        wb.set_balls(W["pos"], W["vel"] * 6.0, W["rad"], W["mass"])
        wb.set_styles([(1, GRAY, BLACK)] * nb, [(0, RED)] * 4, 3, 3)
        return wb

    a, b = make(), make()
    ring = torch.zeros((slots, n, 64, 64, 4), dtype=torch.uint8, device="cuda")
    traj = torch.zeros((steps, n, nb, 2), dtype=torch.float32, device="cuda")
    a.step_ring_async(steps, ring, every, traj=traj)
    torch.cuda.synchronize()
    n_draws = steps // every
    frames = {}
    for s in range(steps):
        t, _ = b.step(1, record_traj=True)
        assert torch.equal(traj[s], t[0]), s
        if (s + 1) % every == 0:
            frames[(s + 1) // every - 1] = b.render().clone()
    for d in range(max(0, n_draws - slots), n_draws):                 # the ring keeps the last `slots` frames
        assert torch.equal(ring[d % slots], frames[d]), d
    pa, va = a.get_balls()
    pb, vb = b.get_balls()
    assert np.array_equal(pa, pb) and np.array_equal(va, vb)
    sa, ka = a.get_status()
    sb, kb = b.get_status()
    assert np.array_equal(sa, sb) and np.array_equal(ka, kb)
    ea, _ = a.read_events()
    eb, _ = b.read_events()
    assert np.array_equal(ea, eb)


def test_full_size_frame_properties():
    """BASELINE configs[3] shard at full size (131,072 worlds x 8 balls, 64x64 frames): properties that
    do not need the oracle -- wall pixels per frame equal the analytic area of the cage rectangles, the
    pixel under every ball centre has the fill colour, rendering is idempotent, balls stay in their cage."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    n, nb = 131072, 8
    wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
    assert wb.sample_rect_worlds(seed=77, **{k: v for k, v in sampler.scaled64_params(nb).items()}) == 0
    wb.set_styles([(1, GRAY, BLACK)] * nb, [(0, RED)] * 4, 3, 3)
    ring = torch.empty((2, n, 64, 64, 4), dtype=torch.uint8, device="cuda")
    wb.step_ring_async(20, ring, 1)
    torch.cuda.synchronize()
    last = ring[(20 - 1) % 2]
    again = wb.render()
    assert torch.equal(again, last)                                   # the pipelined frame == a fresh render
    status, _ = wb.get_status()
    ok = status == 0
    assert ok.mean() > 0.999
    wall = wb.get_walls()
    pos, _ = wb.get_balls()
    # analytic wall area: union of the 4 integer rectangles, rasterised on the host in chunks
    red = ((last[..., 0] == 255) & (last[..., 1] == 0) & (last[..., 2] == 0)).sum(dim=(1, 2)).cpu().numpy()
    X = np.arange(64)[None, None, :] + 0.5
    Y = np.arange(64)[None, :, None] + 0.5
    for c0 in range(0, n, 16384):
        w = wall[c0:c0 + 16384]
        x0, x1 = w[..., 0].min(2), w[..., 0].max(2)               # (chunk, 4)
        y0, y1 = w[..., 1].min(2), w[..., 1].max(2)
        m = np.zeros((len(w), 64, 64), bool)
        for q in range(4):
            m |= (X >= np.round(x0[:, q])[:, None, None]) & (X < np.round(x1[:, q])[:, None, None]) & \
                 (Y >= np.round(y0[:, q])[:, None, None]) & (Y < np.round(y1[:, q])[:, None, None])
        area, got = m.sum((1, 2)), red[c0:c0 + 16384]
        # a ball touching a wall tints a few wall pixels (its stroke ring reaches sThick/2 beyond the
        # physical radius), so the count of pure wall-colour pixels can only fall short, and only slightly
        assert np.all(got <= area) and np.all(area - got <= 8 * 8)
        assert (got == area).mean() > 0.4
    # the pixel containing every ball's (integer-snapped) centre is the fill colour
    cx = np.clip(np.floor(np.abs(pos[..., 0]) + 0.5).astype(np.int64) - 1, 0, 63)   # centre lies on a pixel corner;
    cy = np.clip(np.floor(np.abs(pos[..., 1]) + 0.5).astype(np.int64) - 1, 0, 63)   # the pixel up-left of it is inside
    idx = torch.from_numpy(np.arange(n)[:, None].repeat(nb, 1)).cuda()
    px = last[idx, torch.from_numpy(cy).cuda(), torch.from_numpy(cx).cuda()].cpu().numpy()
    good = (px == np.array([128, 128, 128, 255], np.uint8)).all(-1)
    assert good[ok].mean() > 0.999, good[ok].mean()                   # (a neighbour's ring may cover it in rare contacts)
    lo = wall[:, 0, 0, :][:, None, :] + 3 + 3 - 0.01
    hi = np.stack([wall[:, 1, 0, 0], wall[:, 2, 0, 1]], -1)[:, None, :] - 3 - 3 + 0.01
    assert ((pos >= lo) & (pos <= hi)).all(axis=(1, 2))[ok].mean() > 0.999
    wb.close()


@pytest.mark.parametrize("canvas", [(32, 48), (128, 70), (700, 333), (1000, 37), (2048, 9), (4096, 5), (4100, 6), (68, 64),
                                    (667, 41), (33, 70), (1001, 9), (2047, 4), (4095, 3), (63, 65),
                                    (64, 64), (64, 48), (64, 9), (64, 63), (64, 100), (64, 130)])
def test_frames_band_tiles_identical(canvas):
    """(64-pixel-wide canvases of at most 64 rows: the swizzled tile stored through a tensor map, box = 2 H lines of
    128 bytes; taller ones: square tiles.)  Canvases that are not 64 pixels wide are drawn in bands of whole frame rows (4096 / W rows per band, one
    linear bulk copy each -- or, when W is not a multiple of 4, a stream of 16-byte stores out of a padded buffer;
    wider than 4096 pixels: square tiles again): every band geometry -- several rows, two rows, one row per band,
    a last band that is cut short, balls cut by band edges and by the canvas, small sprites composited in pairs --
    gives the bytes of the restated renderer."""
    W, H = canvas
    rs = np.random.RandomState(W * 7 + H)
    nw, nb = 3, 6
    t = max(2, min(W, H) // 12)
    x1, y1 = W - 3, H - 2
    wall = np.array([[[[2, 1], [x1, 1], [x1, 1 + t], [2, 1 + t]],
                      [[x1, 1], [x1, y1], [x1 - t, y1], [x1 - t, 1]],
                      [[x1, y1], [2, y1], [2, y1 - t], [x1, y1 - t]],
                      [[2, y1], [2, 1], [2 + t, 1], [2 + t, y1]]]] * nw, np.float64)
    pos = np.stack([rs.uniform(-4, W + 4, (nw, nb)), rs.uniform(-4, H + 4, (nw, nb))], -1)
    pos[0, 0] = (W - 0.5, H / 2.0)          # cut by the right edge
    pos[0, 1] = (W / 2.0, -1.0)             # cut by the top edge
    pos[1, :2] = [(10.0, H / 2.0), (21.0, H / 2.0)]   # two small sprites with the same geometry side by side
    for r, st in ((3, 1), (9, 2)):
        dev, ref = _render_both(wall, pos, r, canvas, st)
        assert np.array_equal(dev, ref), (canvas, r)


@pytest.mark.parametrize("canvas", [(64, 64), (200, 96)])
def test_frames_maximum_object_counts(canvas):
    """PPB_MAX_WALLS walls and PPB_MAX_BALLS balls per world: the per-warp metadata is at its largest (fewest warps
    per CTA) and every lane of a warp owns a ball."""
    W, H = canvas
    rs = np.random.RandomState(5)
    nw = 4
    wall = np.zeros((nw, 16, 4, 2), np.float64)
    for w in range(nw):
        for q in range(16):
            x0, y0 = rs.randint(0, W - 8), rs.randint(0, H - 8)
            ww, hh = rs.randint(2, 12), rs.randint(2, 12)
            wall[w, q] = [[x0, y0], [x0 + ww, y0], [x0 + ww, y0 + hh], [x0, y0 + hh]]
    pos = np.stack([rs.uniform(0, W, (nw, 32)), rs.uniform(0, H, (nw, 32))], -1)
    cols = [(rs.rand(), rs.rand(), rs.rand(), 1.0) for _ in range(16)]
    dev, ref = _render_both(wall, pos, 3, canvas, 1, wall_rgba=cols)
    assert np.array_equal(dev, ref)


def test_ticketed_tiles_many_launches_two_batches_two_streams():
    """render_tile_kernel hands its tiles out through ticket counters taken from a ring of 256 pairs that every launch
    re-arms: 600 launches (the ring wraps twice) of two batches interleaved on two streams, each with more frames
    than the grid has warps (so most tiles come from tickets), give the same bytes as the first launch -- and the
    first launch gives the bytes of the restated renderer."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    nb = 8
    batches, want = [], []
    for seed, n in ((5, 9000), (6, 12345)):
        W = sampler.sample_rect_worlds(n, nb, seed=seed, **sampler.scaled64_params(nb))
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64))
        wb.set_walls(W["wall"])
        wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        wb.set_styles([(1, GRAY, BLACK)] * nb, [(0, RED)] * 4, 3, 3)
        first = wb.render().clone()
        ref = oracle.render_frames(W["wall"][:2048], W["pos"][:2048], W["rad"][:2048], (64, 64), s_thick=1, n_threads=8)
        assert np.array_equal(first[:2048].cpu().numpy(), ref)
        tail = oracle.render_frames(W["wall"][-512:], W["pos"][-512:], W["rad"][-512:], (64, 64), s_thick=1, n_threads=8)
        assert np.array_equal(first[-512:].cpu().numpy(), tail)
        batches.append(wb)
        want.append(first)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [torch.empty_like(w) for w in want]
    for it in range(300):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                batches[k].render(out=outs[k])
        if it % 50 == 49 or it < 3:
            torch.cuda.synchronize()
            for k in (0, 1):
                assert torch.equal(outs[k], want[k]), (it, k)
                outs[k].zero_()
    torch.cuda.synchronize()
    for wb in batches:
        wb.close()


def test_frames_follow_radius_changes():
    """The render kernel reads the radii from the per-world render record (wall cache + radii), which is refreshed after
    the balls changed: new radii through set_balls (whole batch, then a partial world range, with empty slots), and radii
    written through the raw device pointer followed by requeue() -- every frame equals the restated pipeline's."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    n, nb = 300, 8
    W = sampler.sample_rect_worlds(n, nb, seed=11, **sampler.scaled64_params(nb))
    wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64))
    wb.set_walls(W["wall"])
    wb.set_styles([(1, GRAY, BLACK)] * nb, [(0, RED)] * 4, 3, 4)
    rng = np.random.RandomState(5)

    def check(rad):
        dev = wb.render().cpu().numpy()
        ref = oracle.render_frames(W["wall"], W["pos"], rad, (64, 64), s_thick=1, n_threads=8)
        assert np.array_equal(dev, ref)

    rad = np.full((n, nb), 3.0)
    wb.set_balls(W["pos"], W["vel"], rad, W["mass"])
    check(rad)
    rad = rng.randint(3, 5, size=(n, nb)).astype(np.float64)          # radii 3 and 4 mixed
    rad[rng.rand(n, nb) < 0.2] = -1.0                                  # empty slots
    wb.set_balls(W["pos"], W["vel"], rad, W["mass"])
    check(rad)
    rad[100:150] = 4.0                                                 # a partial world range
    wb.set_balls(W["pos"][100:150], W["vel"][100:150], rad[100:150], W["mass"][100:150], world0=100)
    check(rad)
    # radii written through the raw pointer take effect after requeue()
    ptr, nbytes = wb.device_ptr(3)
    assert nbytes == n * nb * 2 * 4
    rad = np.where(rad > 0, 3.0, -1.0)                                 # every live ball back to radius 3
    rm = np.stack([rad.astype(np.float32), W["mass"].astype(np.float32)], axis=-1)
    class _Raw:   # zero-copy view of the library's {radius, mass} array (the use ppb_device_ptr documents)
        __cuda_array_interface__ = {"shape": (n, nb, 2), "typestr": "<f4", "data": (ptr, False), "version": 2}
    torch.as_tensor(_Raw(), device="cuda").copy_(torch.from_numpy(np.ascontiguousarray(rm)))
    torch.cuda.synchronize()
    wb.requeue()
    check(rad)
    wb.close()
