"""CPU tests of oracle/cairo_oracle.c, the restatement of the published cairo image pipeline (see its header),
of how far the older exact-area model (oracle/render_oracle.c) is from it, and of the product's per-pixel scan
conversion (csrc/ppb_cairo.cu) evaluated on the host.

libcairo itself is absent (parity with its output is UNPINNED); the known-answer values below are the parameters
cairo's sources state: the arc error table, the 0.1 px flattening tolerance, 24.8 fixed point, 15 sub-rows,
alpha = (17 area + 256) >> 9 with full area 2*256*15.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRAY, BLACK = (0.5, 0.5, 0.5, 1.0), (0.0, 0.0, 0.0, 1.0)

# measured maxima (per channel, of 255) of |cairo restatement - exact-area model|; asserted below
SPRITE_TOL = {(3, 1): 25, (3, 2): 22, (4, 1): 29, (4, 2): 24, (25, 1): 24, (25, 2): 19}
WALL_TOL = 28


def test_arc_segments_follow_the_error_table():
    L = oracle.lib()
    seg = lambda angle, r: L.cairo_arc_segments_needed(angle, float(r), 0.1)
    # one Bezier per half turn while 0.1 / r > 1/54, two down to 2.7e-4, three below that
    assert [seg(np.pi, r) for r in (1, 3, 5, 6, 25, 300, 400)] == [1, 1, 1, 2, 2, 2, 3]


def test_flattened_circle_is_a_closed_polygon_within_the_tolerance():
    for r, n_expected in ((3, 17), (4, 25), (25, 65)):
        c = r + 1
        v = oracle.cairo_circle_flatten(c, c, r)
        assert len(v) == n_expected                     # r = 3: each pi Bezier halved 3 times (16 chords); r = 4: adaptively, 24 chords; r = 25: 64 chords
        assert tuple(v[0]) == tuple(v[-1]) == ((c + r) * 256, c * 256)
        p = v / 256.0 - c
        rad = np.hypot(p[:, 0], p[:, 1])
        assert np.abs(rad - r).max() < 0.02 * r + 1 / 256.0       # Bezier arcs of pi overshoot by <= 1.85 %
        mid = 0.5 * (p[1:] + p[:-1])
        assert (r - np.hypot(mid[:, 0], mid[:, 1])).max() < 0.1   # sagitta of every chord within the tolerance
        ang = np.unwrap(np.arctan2(p[:, 1], p[:, 0]))
        assert np.all(np.diff(ang) > 0) and abs(ang[-1] - ang[0] - 2 * np.pi) < 1e-9


def test_alpha_of_full_empty_and_half_pixels():
    # a 45-degree edge through pixel corners: pixels on the diagonal are half covered, above 255, below 0
    q = np.array([[0.0, 0.0], [8.0, 8.0], [0.0, 8.0], [0.0, 0.0]])
    m = oracle.cairo_quad_mask(q, 8, 8)
    assert np.all(np.tril(m, -1)[np.tril_indices(8, -1)] == 255)
    assert np.all(np.triu(m, 1) == 0)
    # row 0, where the edges start, is sampled at the top of each of its 15 sub-rows: sub-row k covers x in
    # [0, k/15) of the diagonal pixel, area = 2 * sum_k floor(256 k / 15) = 3570 of 7680 -> (17 * 3570 + 256) >> 9
    # = 119; the rows below are stepped whole (no edge starts, ends or crosses): exact trapezoid, 3840 -> 128
    assert m[0, 0] == 119 and set(np.diag(m)[1:]) == {128}
    # axis-aligned boxes go through the rectangular converter: exact area, round half up
    m = oracle.cairo_quad_mask(np.array([[1.5, 1.25], [3.0, 1.25], [3.0, 3.0], [1.5, 3.0]]), 4, 4)
    assert m[2, 2] == 255 and m[1, 1] == round(255 * 0.5 * 0.75) and m[1, 2] == round(255 * 0.75) and m[0, 0] == 0


def test_disc_and_ring_masks():
    for r in (3, 4, 25):
        for st in (1, 2):
            sz = 2 * (r + st)
            d = oracle.cairo_disc_mask(r + st, r + st, r, sz, sz).astype(np.float64) / 255.0
            ring, cusps = oracle.cairo_ring_mask(r + st, r + st, r, st, sz, sz)
            assert cusps == 0                                        # no cusp fan on a circle (not restated)
            ring = ring.astype(np.float64) / 255.0
            # polygonised circle: area within 3 % below pi r^2, never above
            assert 0.97 * np.pi * r * r < d.sum() < 1.005 * np.pi * r * r
            assert abs(ring.sum() - 2 * np.pi * r * st) < 0.04 * 2 * np.pi * r * st
            assert d[r + st, r + st] == 1.0 and d[0, 0] == 0.0 and ring[r + st, r + st] == 0.0


@pytest.mark.parametrize("r,st", sorted(SPRITE_TOL))
def test_area_model_vs_cairo_sprites(r, st):
    """The measured distance between the exact-area sprite (round 1's model) and the cairo restatement."""
    a = oracle.ball_sprite(r, st, GRAY, BLACK, model="cairo").astype(np.int32)
    b = oracle.ball_sprite(r, st, GRAY, BLACK, model="area").astype(np.int32)
    d = np.abs(a - b).max(axis=-1)
    print("sprite r=%d sThick=%d: max |cairo - area| = %d/255, pixels differing %d of %d"
          % (r, st, d.max(), (d > 0).sum(), d.size))
    assert d.max() <= SPRITE_TOL[(r, st)]
    c = a.shape[0] // 2
    assert d[c - 1:c + 1, c - 1:c + 1].max() == 0       # the interior is the fill colour either way


@pytest.mark.parametrize("name,canvas", [("diamond", (640, 480)), ("rhomb3_s300", (1600, 1600)),
                                         ("ngon6_s801", (800, 800)), ("ngon8_s800", (800, 800))])
def test_area_model_vs_cairo_rotated_walls(name, canvas):
    case = [c for c in load_golden("physics_golden.npz") if c.name == name][0]
    kw = dict(order=case["order"], s_thick=2, n_threads=8)
    a = oracle.render_frames(case["wall"][None], case["pos0"][None], case["rad"][None], canvas, **kw).astype(np.int32)
    b = oracle.render_frames(case["wall"][None], case["pos0"][None], case["rad"][None], canvas, rotated_walls="area",
                             **kw).astype(np.int32)
    d = np.abs(a - b).max(axis=-1)
    print("%s: max |cairo - area walls| = %d/255 on %d pixels (%d antialiased wall pixels)"
          % (name, d.max(), (d > 0).sum(), ((b[..., 1] > 0) & (b[..., 1] < 255)).sum()))
    assert d.max() <= WALL_TOL
    assert (d > 0).mean() < 0.01


def test_axis_aligned_frames_agree_between_the_models_away_from_ball_edges():
    from pyphy_engine_b200 import sampler
    W = sampler.sample_rect_worlds(64, 8, seed=2, **sampler.scaled64_params(8))
    a = oracle.render_frames(W["wall"], W["pos"], W["rad"], (64, 64), s_thick=1, model="cairo").astype(np.int32)
    b = oracle.render_frames(W["wall"], W["pos"], W["rad"], (64, 64), s_thick=1, model="area").astype(np.int32)
    d = np.abs(a - b).max(axis=-1)
    assert d.max() <= 30
    flat = (b == np.array([255, 255, 255, 255])).all(-1) | (b == np.array([255, 0, 0, 255])).all(-1) | \
           (b == np.array([128, 128, 128, 255])).all(-1)
    assert (d[flat] > 0).mean() < 0.002                # white, wall red and ball interior are the same pixels


def _premul8(c):
    c8 = lambda v: int(min(max(v, 0.0), 1.0) * 65535.0 + 0.5) >> 8
    return c8(c[0] * c[3]) | (c8(c[1] * c[3]) << 8) | (c8(c[2] * c[3]) << 16) | (c8(c[3]) << 24)


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_product_scan_conversion_on_the_host_equals_the_oracle(tmp_path):
    """csrc/ppb_cairo.cu's per-pixel evaluation (what build_ball_sprites_kernel runs per thread), compiled for the host
    by tools/sprite_host_check.cu: same sprites as the oracle's edge-walking converter, bit for bit."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "sprite_host_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                           os.path.join(ROOT, "tools", "sprite_host_check.cu")])
    styles = [(GRAY, BLACK), ((0.1, 0.7, 0.3, 0.8), (0.3, 0.0, 0.9, 1.0)), ((0.1, 0.7, 0.3, 1.0), (0.3, 0.0, 0.9, 0.4))]
    for fill, stroke in styles:
        for r, st in ((1, 1), (2, 1), (3, 1), (3, 2), (4, 1), (5, 2), (7, 0), (10, 1), (15, 2), (25, 2), (40, 3), (64, 2)):
            out = subprocess.check_output([exe, str(r), str(st), "%x" % _premul8(fill), "%x" % _premul8(stroke)]).split()
            sz = int(out[0])
            dev = np.array([int(x, 16) for x in out[1:]], np.uint32).reshape(sz, sz).view(np.uint8).reshape(sz, sz, 4)
            assert np.array_equal(dev, oracle.ball_sprite(r, st, fill, stroke)), (r, st, fill)
