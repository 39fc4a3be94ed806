"""Consumers of the frame / trajectory tensors: device crops of
DataSaver.fetch(cropSz) and utils.detect_collision, each against a literal numpy restatement of the
reference lines (dataio.py:160-182, utils.py:38-66) written in this test."""
import math

import numpy as np
import pytest
import torch

from pyphy_engine_b200 import sampler
from pyphy_engine_b200.engine import WorldBatch, collision_flags
import pyphy_engine_b200.dataio as dio
import pyphy_engine_b200.utils as ut

pytestmark = pytest.mark.gpu


def py2_round(v):
    return math.floor(abs(v) + 0.5) * (-1.0 if v < 0 else 1.0)


def ref_crop(im, px, py, cropSz, xSz, ySz):
    """dataio.py:160-179 for one ball in one frame."""
    xMid, yMid = py2_round(px), py2_round(py)
    out = 255 * np.ones((cropSz, cropSz, 3), np.uint8)
    x1, x2 = max(0, xMid - cropSz / 2.0), min(xSz, xMid + cropSz / 2.0)
    y1, y2 = max(0, yMid - cropSz / 2.0), min(ySz, yMid + cropSz / 2.0)
    imX1, imX2 = int(py2_round(cropSz / 2.0 - (xMid - x1))), int(py2_round(cropSz / 2.0 + (x2 - xMid)))
    imY1, imY2 = int(py2_round(cropSz / 2.0 - (yMid - y1))), int(py2_round(cropSz / 2.0 + (y2 - yMid)))
    x1, x2, y1, y2 = int(py2_round(x1)), int(py2_round(x2)), int(py2_round(y1)), int(py2_round(y2))
    out[imY1:imY2, imX1:imX2, :] = im[y1:y2, x1:x2, 0:3]
    pos = np.float32(np.float32(np.float32(px) - x1) + imX1), np.float32(np.float32(np.float32(py) - y1) + imY1)
    return out, pos


@pytest.mark.parametrize("dtype,crop", [("f32", 16), ("f64", 24), ("f32", 40)])
def test_device_crops_match_the_reference_slicing(dtype, crop):
    n, nb, steps = 96, 4, 12
    W = sampler.sample_rect_worlds(n, nb, seed=4, **sampler.scaled64_params(nb))
    wb = WorldBatch(n, nb, 4, dtype=dtype, canvas=(64, 64))
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"] * 15.0, W["rad"], W["mass"])       # fast balls: some reach the borders
    wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
    traj, frames = wb.step(steps, record_traj=True, render_every=1)
    crops, cpos = wb.crop(frames, traj, crop)
    crops, cpos = crops.cpu().numpy(), cpos.cpu().numpy()
    fr, tr = frames.cpu().numpy(), traj.cpu().numpy().astype(np.float64)
    for s in (0, steps - 1):
        for w in range(0, n, 7):
            for b in range(nb):
                ref, rpos = ref_crop(fr[s, w], tr[s, w, b, 0], tr[s, w, b, 1], crop, 64, 64)
                assert np.array_equal(crops[s, w, b], ref)
                assert (cpos[s, w, b, 0], cpos[s, w, b, 1]) == rpos
    assert (crops != 255).any()
    wb.close()


def test_fetch_with_crops_through_datasaver(tmp_path):
    kw = dict(wThick=3, isRect=True, mnBallSz=3, mxBallSz=4, arenaSz=64, mnWLen=36, mxWLen=50, numBalls=2,
              mnForce=5.0, mxForce=12.0, mnSeqLen=20, mxSeqLen=30)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    ims = ds.fetch()
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    imBalls, force, velocity, position = ds.fetch(cropSz=16)
    L = len(ims)
    assert len(imBalls) == 2 and len(imBalls[0]) == L and imBalls[0][0].shape == (16, 16, 3)
    # the crop of frame i equals the reference slicing of the full frame i around the recorded centre
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    specs, pos_full, _ = ds.fetch_batch(1, want_frames=False)
    for i in (0, L // 2, L - 1):
        for j in range(2):
            px, py = float(pos_full[0][2 * j, i]), float(pos_full[0][2 * j + 1, i])
            # the recorded float32 centre rounds like the fp64 one unless it sits within 1e-5 of a half pixel
            if abs((px % 1) - 0.5) < 1e-4 or abs((py % 1) - 0.5) < 1e-4:
                continue
            ref, rpos = ref_crop(ims[i], px, py, 16, 64, 64)
            assert np.array_equal(imBalls[j][i], ref)
            assert abs(position[2 * j, i] - rpos[0]) < 1e-4 and abs(position[2 * j + 1, i] - rpos[1]) < 1e-4
    assert np.count_nonzero(velocity) == 0                      # filled only with procSz (dataio.py:184-189)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=5, makeDirs=False, **kw)
    imB8, _, vel8, pos8 = ds.fetch(cropSz=16, procSz=8)
    assert imB8[0][0].shape == (8, 8, 3) and np.allclose(pos8, position * 0.5, atol=1e-5)
    assert np.count_nonzero(vel8[:, :-1]) > 0 and np.count_nonzero(vel8[:, -1]) == 0


def ref_detect(pos, xMin, xMax, yMin, yMax, thresh=0.03):
    """utils.py:38-66 for the (2, L) float32 positions of one ball."""
    velB = np.diff(pos)
    accB = np.diff(velB)
    accB = np.sum(accB * accB, axis=0)
    velB = np.sum(velB * velB, axis=0)
    idx = np.where(accB >= thresh * velB[0:-1])[0] + 1
    fIdx, bCol = [], []
    for i in range(len(idx)):
        if idx[i] in fIdx:
            continue
        if i < len(idx) - 1 and (idx[i] + 1 == idx[i + 1]):
            fIdx.append(int(idx[i + 1]))
        else:
            fIdx.append(int(idx[i]))
    for i in fIdx:
        p = min(i, pos.shape[1] - 1)
        x, y = pos[:, p]
        bCol.append(bool(x > xMin and x < xMax and y > yMin and y < yMax))
    return fIdx, bCol


def test_detect_collision_matches_the_reference_heuristic(tmp_path):
    kw = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=700, mnWLen=300, mxWLen=550, numBalls=3,
              mnForce=3e4, mxForce=8e4, mnSeqLen=150, mxSeqLen=150)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=1094, makeDirs=False, **kw)
    specs, position, _ = ds.fetch_batch(6, want_frames=False)
    n_col = 0
    for sp, pos in zip(specs, position):
        cage = np.array([[p.x(), p.y()] for p in sp.cage_pts])
        allfIdx, allbCol = ut.detect_collision_arrays(pos, cage)
        xMin, yMin = cage.min(0) + 65
        xMax, yMax = cage.max(0) - 65
        for b in range(3):
            fIdx, bCol = ref_detect(pos[2 * b:2 * b + 2], xMin, xMax, yMin, yMax)
            assert allfIdx[b] == fIdx and allbCol[b] == bCol
            n_col += len(fIdx)
    assert n_col > 10
    # the raw flags on a big random tensor, bit for bit against numpy float32 arithmetic
    rng = np.random.RandomState(0)
    traj = np.cumsum(rng.randn(50, 5000, 2).astype(np.float32), axis=0).astype(np.float32)
    flags = collision_flags(torch.from_numpy(traj).cuda(), 0.03).cpu().numpy()
    v = np.diff(traj, axis=0)
    a = np.diff(v, axis=0)
    ref = (np.sum(a * a, -1) >= np.float32(0.03) * np.sum(v * v, -1)[:-1]).T
    assert np.array_equal(flags.astype(bool), ref)


def test_detect_collision_from_a_saved_tree(tmp_path):
    kw = dict(wThick=30, isRect=True, mnBallSz=25, mxBallSz=25, arenaSz=700, mnWLen=300, mxWLen=550, numBalls=2,
              mnForce=8e4, mxForce=8e4, mnSeqLen=40, mxSeqLen=40)
    ds = dio.DataSaver(rootPath=str(tmp_path), randSeed=4185, **kw)
    ds.save(numSeq=2)
    allfIdx, allbCol = ut.detect_collision(ds.seqDir_ % 0)
    assert len(allfIdx) == 2 and all(len(a) == len(b) for a, b in zip(allfIdx, allbCol))
    col, bcol = ut.det_collisions_all(ds.dirName_)
    assert col >= sum(len(a) for a in allfIdx) and 0 <= bcol <= col
    n = ut.save_collisions_seq(ds.dirName_, isSubFolder=False, outFolder=str(tmp_path / "clips"))
    assert n == col


def test_custom_world_glimpse(tmp_path):
    """CustomWorld.get_glimpse (dataio.py:490-519) against a literal restatement of its cropping lines."""
    import pyphy_engine_b200.geometry as gm
    cw = dio.CustomWorld(numballs=2, ballsz=20, wthick=20, wtheta=[0, 90, -180], wlen=300, arenasz=400,
                         ballLocs=[gm.Point(120, 120), gm.Point(250, 200)],
                         forces=[gm.Point(5e4, 3e4), gm.Point(-4e4, 2e4)])
    cw.generate_model()
    assert [n for n in cw.world_.get_object_names()] == ["wall-0", "wall-1", "wall-2", "wall-3", "ball-0", "ball-1"]
    for it in range(25):
        imBalls, position, velocity = cw.get_glimpse(cropSz=96)
        im = cw.model_.generate_image()
        assert len(imBalls) == 2 and imBalls[0].shape == (96, 96, 3)
        for b in range(2):
            px, py = position[2 * b], position[2 * b + 1]
            xMid, yMid = py2_round(px), py2_round(py)
            ref = 255 * np.ones((96, 96, 3), np.uint8)
            x1, x2 = max(0, xMid - 48.0), min(400, xMid + 48.0)
            y1, y2 = max(0, yMid - 48.0), min(400, yMid + 48.0)
            imX1, imX2 = int(48.0 - (xMid - x1)), int(48.0 + (x2 - xMid))
            imY1, imY2 = int(48.0 - (yMid - y1)), int(48.0 + (y2 - yMid))
            ref[imY1:imY2, imX1:imX2, :] = im[int(y1):int(y2), int(x1):int(x2), 0:3]
            assert np.array_equal(imBalls[b], ref)
    assert not np.allclose(position, [120, 120, 250, 200])           # the balls moved
    assert dio.fix_theta_range(270) == -90 and dio.fix_theta_range(-180) == 180 and dio.fix_theta_range(90) == 90


@pytest.mark.parametrize("dtype,crop,proc", [("f32", 16, 8), ("f64", 24, 10), ("f32", 40, 64), ("f32", 17, 5), ("f32", 32, 32),
                                              ("f32", 64, 7)])
def test_device_resize_equals_pil_bilinear(dtype, crop, proc):
    """fetch(cropSz, procSz) (dataio.py:183-189): scipy.misc.imresize(imBall, (procSz, procSz)) is
    PIL.Image.resize(..., BILINEAR) of the uint8 crop; the device kernel restates Pillow's 8-bit resampler and is
    compared here with Pillow itself (an implementation this repository did not write), byte for byte.  Positions
    and displacements follow the reference's expressions literally."""
    from PIL import Image
    n, nb, steps = 24, 3, 6
    W = sampler.sample_rect_worlds(n, nb, seed=9, **sampler.scaled64_params(4))
    wb = WorldBatch(n, nb, 4, dtype=dtype, canvas=(64, 64))
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"][:, :nb], W["vel"][:, :nb] * 15.0, W["rad"][:, :nb], W["mass"][:, :nb])
    wb.set_styles([(1, (.5, .5, .5, 1), (0, 0, 0, 1))] * nb, [(0, (1, 0, 0, 1))] * 4, 3, 3)
    traj, frames = wb.step(steps, record_traj=True, render_every=1)
    crops, cpos = wb.crop(frames, traj, crop)
    crops_h, cpos_h = crops.cpu().numpy(), cpos.cpu().numpy().copy()
    out, pos, vel = wb.resize_crops(crops, proc, pos=cpos, traj=traj, want_vel=True)
    out, pos, vel = out.cpu().numpy(), pos.cpu().numpy(), vel.cpu().numpy()
    tr = traj.cpu().numpy().astype(np.float64)
    scale = float(proc) / float(crop)
    for s in range(steps):
        for w in range(0, n, 5):
            for b in range(nb):
                ref = np.asarray(Image.fromarray(crops_h[s, w, b]).resize((proc, proc), Image.BILINEAR))
                assert np.array_equal(out[s, w, b], ref), (s, w, b)
                # position[2j, i] = position[2j, i] * posScale on a float32 array element (dataio.py:186-187)
                assert pos[s, w, b, 0] == np.float32(np.float64(cpos_h[s, w, b, 0]) * scale)
                assert pos[s, w, b, 1] == np.float32(np.float64(cpos_h[s, w, b, 1]) * scale)
                if s > 0:       # velocity[2j, i-1] = (pos.x() - pPos) * posScale (dataio.py:151-152, 189-190)
                    assert vel[s - 1, w, b, 0] == np.float32((tr[s, w, b, 0] - tr[s - 1, w, b, 0]) * scale)
                    assert vel[s - 1, w, b, 1] == np.float32((tr[s, w, b, 1] - tr[s - 1, w, b, 1]) * scale)
    assert not vel[steps - 1].any()
    # random content as well: every coefficient pattern, values over the whole byte range
    rs = np.random.RandomState(crop * 100 + proc)
    noise = torch.from_numpy(rs.randint(0, 256, (steps, n, nb, crop, crop, 3)).astype(np.uint8)).cuda()
    out2, _, _ = wb.resize_crops(noise, proc)
    out2, noise = out2.cpu().numpy(), noise.cpu().numpy()
    for idx in ((0, 0, 0), (2, 7, 1), (5, 23, 2)):
        ref = np.asarray(Image.fromarray(noise[idx]).resize((proc, proc), Image.BILINEAR))
        assert np.array_equal(out2[idx], ref), idx
    wb.close()


@pytest.mark.parametrize("dtype", ["f32", "f64"])
def test_horizon_matrices_of_a_recorded_trajectory(dtype):
    """DynamicsHorizon.get_data's matrix (primitives.py:843-850) for all frames in one launch, against the literal
    loop: outMat[i, 2n], outMat[i, 2n+1] = float32(pos[n+1] - pos[n]) over the look-ahead window."""
    n, nb, steps, N = 50, 4, 30, 7
    W = sampler.sample_rect_worlds(n, nb, seed=6, **sampler.scaled64_params(nb))
    wb = WorldBatch(n, nb, 4, dtype=dtype)
    wb.set_walls(W["wall"])
    wb.set_balls(W["pos"], W["vel"] * 10.0, W["rad"], W["mass"])
    p0, _ = wb.get_balls()
    traj, _ = wb.step(steps, record_traj=True)
    full = torch.cat([torch.from_numpy(p0.astype(np.float64 if dtype == "f64" else np.float32))[None].cuda(), traj], 0)
    mats = wb.horizon(full.contiguous(), N).cpu().numpy()
    assert mats.shape == (steps + 1 - N, n, nb, 2 * N) and mats.dtype == np.float32
    rows = full.cpu().numpy()
    for t in (0, 5, steps - N):
        for w in (0, 17, 49):
            for b in range(nb):
                for k in range(N):
                    d = rows[t + k + 1, w, b] - rows[t + k, w, b]
                    assert mats[t, w, b, 2 * k] == np.float32(d[0]) and mats[t, w, b, 2 * k + 1] == np.float32(d[1])
    wb.close()
