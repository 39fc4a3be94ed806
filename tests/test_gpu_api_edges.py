"""C-ABI edge cases and error behaviour on a real device: argument validation, maximum slot counts,
empty slots, tiny and ragged batches, state-machine errors."""
import ctypes

import numpy as np
import pytest

import oracle
from pyphy_engine_b200 import _capi, sampler
from pyphy_engine_b200.engine import WorldBatch

pytestmark = pytest.mark.gpu


def test_argument_validation():
    with pytest.raises(_capi.PpbError, match="n_balls"):
        WorldBatch(4, 0, 4)
    with pytest.raises(_capi.PpbError, match="n_balls"):
        WorldBatch(4, 33, 4)
    with pytest.raises(_capi.PpbError, match="n_walls"):
        WorldBatch(4, 2, 17)
    with pytest.raises(_capi.PpbError, match="dt"):
        WorldBatch(4, 2, 4, dt=0.0)
    with pytest.raises(_capi.PpbError, match="permutation"):
        WorldBatch(4, 2, 2, order=[0, 0, 2, 3])
    with pytest.raises(_capi.PpbError, match="increasing"):
        WorldBatch(4, 2, 1, order=[0, 2, 1])
    with pytest.raises(_capi.PpbError, match="device"):
        WorldBatch(4, 2, 4, device=99)
    wb = WorldBatch(4, 2, 4)
    with pytest.raises(_capi.PpbError, match="range"):
        wb.get_balls(world0=3, count=5)
    with pytest.raises(_capi.PpbError, match="mass"):
        wb.set_balls(np.zeros((4, 2, 2)), np.zeros((4, 2, 2)), np.ones((4, 2)), np.zeros((4, 2)))
    wc = WorldBatch(4, 2, 4, canvas=(32, 32))
    with pytest.raises(_capi.PpbError, match="ppb_set_styles"):
        wc.render()
    with pytest.raises(_capi.PpbError, match="ppb_set_styles"):
        wc.step(2, render_every=1)
    wc.close()
    import torch
    wb.set_styles([(1, (0, 0, 0, 1), (0, 0, 0, 1))] * 2, [(0, (1, 0, 0, 1))] * 4, 3, 3)
    with pytest.raises(_capi.PpbError, match="canvas"):
        wb.render(out=torch.empty((4, 8, 8, 4), dtype=torch.uint8, device="cuda"))
    lib = _capi.load()
    assert lib.ppb_step_async(None, 1, None, None, 1, None) == -1 and b"null batch" in lib.ppb_last_error()
    assert lib.ppb_version() >= 1
    wb.close()
    wb.close()      # idempotent


def test_max_walls_and_empty_slots_bit_exact():
    """16 wall slots (64 collision edges: a 16-gon cage) and ball slots left empty (radius < 0)."""
    rng = np.random.RandomState(3)
    n, nb, nw = 300, 6, 16
    ang = 2 * np.pi * np.arange(nw) / nw
    R = 300.0
    pts = np.stack([400 + R * np.cos(ang), 400 + R * np.sin(ang)], -1)            # clockwise on screen (y down)
    wall = sampler.cage_vertices(np.broadcast_to(pts, (n, nw, 2)).copy(), 20.0)
    pos = np.zeros((n, nb, 2)); vel = rng.uniform(-900, 900, (n, nb, 2))
    rad = np.full((n, nb), 18.0)
    for w in range(n):                                                              # non-overlapping starts on a small ring
        a0 = rng.uniform(0, 2 * np.pi)
        for b in range(nb):
            pos[w, b] = (400 + 120 * np.cos(a0 + b * 2 * np.pi / nb), 400 + 120 * np.sin(a0 + b * 2 * np.pi / nb))
    rad[::3, 4:] = -1.0                                                             # every third world has 2 empty slots
    mass = sampler.ball_mass(np.abs(rad))
    wb = WorldBatch(n, nb, nw, dtype="f64", event_capacity=1 << 20, max_substeps=4096)
    wb.set_walls(wall)
    wb.set_balls(pos, vel, rad, mass)
    ob = oracle.OracleBatch(wall, pos, vel, rad, mass, trig_mode=1, max_substeps=4096)
    traj, _ = wb.step(120, record_traj=True)
    otraj, oev = ob.step(120, record_traj=True, max_events=1 << 20)
    status, _ = wb.get_status()
    assert np.array_equal(status, ob.status)
    alive = ob.status == 0
    live_slots = (rad >= 0)[alive]
    assert np.array_equal(traj.cpu().numpy()[:, alive][:, live_slots], otraj[:, alive][:, live_slots])
    ev, total = wb.read_events()
    assert total == len(oev) > 0 and np.array_equal(ev, oev)
    assert alive.mean() > 0.9
    wb.close()


def test_single_world_single_ball_and_ragged_sizes():
    for n, nb in ((1, 1), (33, 5), (257, 7)):
        W = sampler.sample_rect_worlds(n, nb, seed=n, arena=1200, mn_wlen=700, mx_wlen=1000)
        wb = WorldBatch(n, nb, 4, dtype="f64", event_capacity=1 << 18)
        wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        ob = oracle.OracleBatch(W["wall"], W["pos"], W["vel"], W["rad"], W["mass"], trig_mode=1)
        traj, _ = wb.step(80, record_traj=True)
        otraj, oev = ob.step(80, record_traj=True, max_events=1 << 18)
        ok = ob.status == 0
        assert np.array_equal(wb.get_status()[0], ob.status)
        assert np.array_equal(traj.cpu().numpy()[:, ok], otraj[:, ok])
        assert np.array_equal(wb.read_events()[0], oev)
        wb.close()


def test_event_buffer_overflow_is_counted_not_fatal():
    W = sampler.sample_rect_worlds(2048, 4, seed=9)
    wb = WorldBatch(2048, 4, 4, dtype="f32", event_capacity=64, max_substeps=256)
    wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.step(200)
    ev, total = wb.read_events()
    assert len(ev) == 64 and total > 1000
    assert wb.counters()["events"] == total
    wb.clear_events()
    assert wb.read_events()[1] == 0
    wb.close()


def test_partial_world_range_updates():
    """set_walls / set_balls / requeue on a sub-range leave the other worlds untouched."""
    W = sampler.sample_rect_worlds(64, 3, seed=2)
    a = WorldBatch(64, 3, 4, dtype="f64"); b = WorldBatch(64, 3, 4, dtype="f64")
    for wb in (a, b):
        wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        wb.step(30)
    # rewrite worlds 16..31 of `a` with their initial state, step both, compare outside / inside the range
    a.set_walls(W["wall"][16:32], world0=16)
    a.set_balls(W["pos"][16:32], W["vel"][16:32], W["rad"][16:32], W["mass"][16:32], world0=16)
    a.step(30); b.step(30)
    pa, _ = a.get_balls(); pb, _ = b.get_balls()
    keep = np.r_[0:16, 32:64]
    assert np.array_equal(pa[keep], pb[keep])
    c = WorldBatch(16, 3, 4, dtype="f64")
    c.set_walls(W["wall"][16:32]); c.set_balls(W["pos"][16:32], W["vel"][16:32], W["rad"][16:32], W["mass"][16:32])
    c.step(30)
    assert np.array_equal(pa[16:32], c.get_balls()[0])
    for wb in (a, b, c):
        wb.close()


def test_simulate_host_async_double_buffered():
    """ppb_simulate_host_async + ppb_host_sync(keep_in_flight): the pipelined host-buffer loop (upload of batch
    i+1 while batch i downloads) returns the same bytes as the blocking call, batch after batch."""
    import torch
    n, nb = 512, 4
    sets = [sampler.sample_rect_worlds(n, nb, seed=40 + k, **sampler.scaled64_params(nb)) for k in range(5)]
    styles = ([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)

    def make():
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64))
        wb.set_walls(sets[0]["wall"]); wb.set_balls(sets[0]["pos"], sets[0]["vel"], sets[0]["rad"], sets[0]["mass"])
        wb.set_styles(*styles)
        return wb

    ref_t, ref_f = [], []
    wb = make()
    for W in sets:
        wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        t = np.empty((3, n, nb, 2), np.float32); f = np.empty((3, n, 64, 64, 4), np.uint8)
        wb.simulate_host(3, t, f, 1)
        ref_t.append(t); ref_f.append(f)
    wb.close()

    wb = make()
    tb = [torch.empty((3, n, nb, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    fb = [torch.empty((3, n, 64, 64, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i, W in enumerate(sets):
        wb.set_walls(W["wall"]); wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
        wb.simulate_host(3, tb[i & 1], fb[i & 1], 1, wait=False)
        wb.host_sync(keep_in_flight=1)
        if i > 0:   # batch i-1 is complete while batch i is still in flight
            assert np.array_equal(tb[(i - 1) & 1].numpy(), ref_t[i - 1])
            assert np.array_equal(fb[(i - 1) & 1].numpy(), ref_f[i - 1])
    wb.host_sync()
    assert np.array_equal(tb[(len(sets) - 1) & 1].numpy(), ref_t[-1])
    assert np.array_equal(fb[(len(sets) - 1) & 1].numpy(), ref_f[-1])
    assert _capi.load().ppb_host_sync(wb._h, 2, None) != 0      # keep_in_flight must be 0 or 1
    wb.close()


def _ref_apply_force(vel, mass, force, force_t):
    """Dynamics.apply_force restated literally (primitives.py:580-594): a = (1.0/mass)*F;
    deltaV = ((0.5*forceT)*forceT)*a; vel = vel + deltaV -- Python floats, one rounding per operation."""
    out = np.empty_like(vel)
    for idx in np.ndindex(vel.shape[:-1]):
        m = float(mass[idx])
        for c in range(2):
            a = (1.0 / m) * float(force[idx + (c,)])
            out[idx + (c,)] = float(vel[idx + (c,)]) + a * (0.5 * force_t * force_t)
    return out


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_apply_force_matches_the_reference_formula(dtype):
    """ppb_apply_force (SURVEY 8b export list): bit-identical with the reference's operation order in fp64, the
    fp64 result rounded to float in fp32 batches; empty slots and the collision caches stay untouched (Q7)."""
    rng = np.random.RandomState(5)
    n, nb = 257, 3
    W = sampler.sample_rect_worlds(n, nb, seed=9, mn_ball=15, mx_ball=35)
    rad = W["rad"].copy()
    rad[::7, 1] = -1.0                                    # some empty slots
    wb = WorldBatch(n, nb, 4, dtype=dtype)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], rad, W["mass"])
    wb.step(3)
    tc0, pa0 = wb.get_collision_cache()
    _, v0 = wb.get_balls()
    F = rng.uniform(-8e4, 8e4, (n, nb, 2))
    wb.apply_force(F, 0.7)
    _, v1 = wb.get_balls()
    # an fp32 batch holds its masses in fp32; the formula itself is evaluated in fp64 and the result rounded once
    mass = W["mass"] if dtype == "f64" else W["mass"].astype(np.float32).astype(np.float64)
    want = _ref_apply_force(v0, mass, F, 0.7)
    if dtype == "f32":
        want = want.astype(np.float32).astype(np.float64)
    occupied = rad >= 0
    assert np.array_equal(v1[occupied], want[occupied])
    assert np.array_equal(v1[~occupied], v0[~occupied])
    tc1, pa1 = wb.get_collision_cache()
    assert np.array_equal(tc0, tc1) and np.array_equal(pa0, pa1)
    # a sub-range of worlds
    wb.apply_force(F[10:20], 1.0, world0=10)
    _, v2 = wb.get_balls()
    assert np.array_equal(v2[:10], v1[:10]) and np.array_equal(v2[20:], v1[20:])
    assert not np.array_equal(v2[10:20], v1[10:20])
    with pytest.raises(_capi.PpbError, match="range"):
        wb.apply_force(F, 1.0, world0=5)
    wb.close()


def test_apply_force_then_step_matches_datasaver_golden():
    """DataSaver kicks every ball once before the first step (dataio.py:382-416): worlds built from the golden's
    ball positions with zero velocity + ppb_apply_force(force, 1.0) step exactly like the recorded reference run."""
    from conftest import load_golden
    for case in [c for c in load_golden("physics64_golden.npz")][:6]:
        nB = case["pos0"].shape[0]
        wb = WorldBatch(1, nB, 4, dtype="f64", event_capacity=1 << 16)
        wb.set_walls(case["wall"][None])
        wb.set_balls(case["ball_pos"][None], np.zeros((1, nB, 2)), case["rad"][None], case["mass"][None])
        wb.apply_force(case["force"][None], 1.0)
        _, v = wb.get_balls()
        assert np.array_equal(v[0], case["vel0"]), case.name
        traj, _ = wb.step(40, record_traj=True)
        assert np.max(np.abs(traj.cpu().numpy()[:, 0] - case["traj"][:40])) < 1e-6
        wb.close()


def test_frames_refuse_radii_without_a_sprite():
    """A radius outside [r_min, r_max] of ppb_set_styles, or a non-integer one, has no pre-rendered sprite
    (get_ball_im uses the radius as an array shape, primitives.py:39-40): every call that draws frames fails
    with PPB_E_STATE instead of silently dropping or truncating the ball."""
    import torch
    W = sampler.sample_rect_worlds(64, 2, seed=4, **sampler.scaled64_params(2))
    style = ([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * 2, [(0, (1, 0, 0, 1))] * 4)
    wb = WorldBatch(64, 2, 4, dtype="f32", canvas=(64, 64))
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wb.set_styles(*style, 4, 6)                                    # radii are 3
    with pytest.raises(_capi.PpbError, match="sprites for"):
        wb.render()
    with pytest.raises(_capi.PpbError, match="sprites for"):
        wb.step(2, render_every=1)
    host = torch.empty((1, 64, 64, 64, 4), dtype=torch.uint8).pin_memory()
    with pytest.raises(_capi.PpbError, match="sprites for"):
        wb.simulate_host(1, None, host, 1)
    wb.step(2)                                                     # physics alone is unaffected
    wb.set_styles(*style, 3, 3)
    wb.render()                                                    # now fine
    rad = W["rad"].copy()
    rad[5, 1] = 3.5
    wb.set_balls(W["pos"], W["vel"], rad, W["mass"])
    with pytest.raises(_capi.PpbError, match="not an integer"):
        wb.render()
    rad[5, 1] = -1.0                                               # an empty slot is not a radius
    wb.set_balls(W["pos"], W["vel"], rad, W["mass"])
    wb.render()
    wb.close()
    # no canvas: ppb_simulate_host refuses frames like ppb_step_ring_async does
    wn = WorldBatch(64, 2, 4, dtype="f32")
    wn.set_walls(W["wall"])
    wn.set_balls(W["pos"], W["vel"], W["rad"], W["mass"])
    wn.set_styles(*style, 3, 3)
    with pytest.raises(_capi.PpbError, match="canvas"):
        wn.simulate_host(1, None, host, 1)
    wn.close()


def test_packed_frame_download_is_lossless():
    """ppb_download_frames / ppb_simulate_host*: frames cross PCIe as run-length tokens and are rebuilt on the host.
    Whatever the content -- rendered canvases, runs of every length, runs ending at the 64-pixel group borders, noise
    that does not compress, pixels that are not opaque (both travel raw) -- the host array equals the device array."""
    import torch
    from pyphy_engine_b200 import sampler
    from pyphy_engine_b200.engine import WorldBatch
    for canvas in ((64, 64), (100, 37)):
        Wd, Hd = canvas
        n, nb = 700, 4
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=canvas)
        rs = np.random.RandomState(Wd)
        npx = Wd * Hd
        cases = {}
        # runs of every length 1..200 cycling through a few colours, starting at a random phase per frame
        lens = np.concatenate([np.arange(1, 201), rs.randint(1, 9, 4000)])
        base = np.repeat(rs.randint(0, 1 << 24, len(lens)).astype(np.uint32) | np.uint32(0xff000000), lens)
        a = np.stack([np.roll(base, int(rs.randint(0, len(base))))[:npx] for _ in range(n)])
        cases["runs"] = a
        cases["flat"] = np.full((n, npx), 0xffffffff, np.uint32)
        noise = rs.randint(0, 1 << 24, (n, npx)).astype(np.uint32) | np.uint32(0xff000000)
        cases["noise"] = noise                               # does not compress: raw
        alpha = a.copy()
        alpha[5, 17] &= np.uint32(0x80ffffff)                # one translucent pixel: raw
        cases["alpha"] = alpha
        half = a.copy()
        half[::2] = noise[::2]                               # every other frame is noise: still below half the raw size?
        cases["mixed"] = half
        for name, arr in cases.items():
            dev = torch.from_numpy(arr.view(np.uint8).reshape(n, Hd, Wd, 4).copy()).cuda()
            got = wb.download_frames(dev)
            assert np.array_equal(got, dev.cpu().numpy()), (canvas, name)
            pinned = torch.empty((n, Hd, Wd, 4), dtype=torch.uint8).pin_memory()
            wb.download_frames(dev, out=pinned)
            assert torch.equal(pinned, dev.cpu()), (canvas, name)
        wb.close()
    # the stepping path: frames of ppb_simulate_host equal the frames rendered on the device, chunk after chunk
    n, nb = 5000, 8
    W = sampler.sample_rect_worlds(n, nb, seed=4, **sampler.scaled64_params(nb))

    def make():
        wb = WorldBatch(n, nb, 4, dtype="f32", canvas=(64, 64), max_substeps=64)
        wb.set_walls(W["wall"])
        wb.set_balls(W["pos"], W["vel"] * 8.0, W["rad"], W["mass"])
        wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
        return wb
    a, b = make(), make()
    host = np.zeros((20, n, 64, 64, 4), np.uint8)
    traj = np.zeros((60, n, nb, 2), np.float32)
    a.simulate_host(60, traj, host, 3)
    t, f = b.step(60, record_traj=True, render_every=3)
    assert np.array_equal(host, f.cpu().numpy())
    assert np.array_equal(traj, t.cpu().numpy())
    a.close()
    b.close()
