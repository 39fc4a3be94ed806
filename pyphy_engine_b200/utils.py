"""Dataset post-processing with the reference's ``utils.py`` interface (utils.py:11-131).

``detect_collision(folderName)`` infers collision frames from a saved sequence's ``position``
array -- a frame is a collision when the squared second difference of a ball's position is at
least 3 % of its squared first difference (utils.py:38-49) -- merges detections in consecutive
frames and labels each one ball-ball (inside the cage shrunk by ``wThick + ballSz + 10``) or
ball-wall.  The per-frame test runs on the GPU over the whole trajectory tensor
(``ppb_collision_flags_async``); the reference's little merge loop stays on the host.
``detect_collision_arrays`` is the same on arrays, for trajectories that never touched the disk.
"""
from __future__ import annotations

import os
import pickle
import shutil
from os import path as osp

import numpy as np

__all__ = ["detect_collision", "detect_collision_arrays", "det_collisions_all", "save_collisions_seq"]


def _merge(idx):
    """utils.py:50-58: of two detections in consecutive frames keep the later one."""
    out = []
    for i in range(len(idx)):
        if idx[i] in out:
            continue
        if i < len(idx) - 1 and idx[i] + 1 == idx[i + 1]:
            out.append(int(idx[i + 1]))
        else:
            out.append(int(idx[i]))
    return out


def detect_collision_arrays(position, cage_xy, wThick=30, ballSz=25, thresh=0.03, device=0):
    """position: (2*nb, seqLen) float32 as in data.mat; cage_xy: (4, 2) cage corner points.
    Returns (allfIdx, allbCol): per ball the collision frames and whether each is ball-ball."""
    import torch
    from .engine import collision_flags
    position = np.ascontiguousarray(position, np.float32)
    nb, L = position.shape[0] // 2, position.shape[1]
    cage_xy = np.asarray(cage_xy, np.float64)
    margin = 10
    xMin, yMin = cage_xy.min(0) + (wThick + ballSz + margin)
    xMax, yMax = cage_xy.max(0) - (wThick + ballSz + margin)
    allfIdx, allbCol = [], []
    if L >= 3:
        traj = torch.from_numpy(position.T.reshape(L, nb, 2).copy()).to(torch.device("cuda", device))
        flags = collision_flags(traj, thresh).cpu().numpy()          # (nb, L - 2)
    else:
        flags = np.zeros((nb, 0), np.uint8)
    for b in range(nb):
        fIdx = _merge(np.nonzero(flags[b])[0] + 1)
        bCol = []
        for i in fIdx:
            p = min(i, L - 1)
            x, y = position[2 * b, p], position[2 * b + 1, p]
            bCol.append(bool(xMin < x < xMax and yMin < y < yMax))
        allfIdx.append(fIdx)
        allbCol.append(bCol)
    return allfIdx, allbCol


def detect_collision(folderName, wThick=30, ballSz=25):
    import scipy.io as sio
    with open(osp.join(folderName, "world.pkl"), "rb") as fh:
        world = pickle.load(fh)
    cage = np.array([[p.x(), p.y()] for p in world["walls"][:4]])
    pos = sio.loadmat(osp.join(folderName, "data.mat"))["position"]
    return detect_collision_arrays(pos, cage, wThick=wThick, ballSz=ballSz)


def det_collisions_all(folderName, limit=1000):
    """(collisions, ball-ball collisions) over the first ``limit`` sequences of a dataset (utils.py:72-85)."""
    names = sorted(f for f in os.listdir(folderName) if osp.isdir(osp.join(folderName, f)))
    col = bcol = 0
    for f in names[:limit]:
        allfIdx, allbCol = detect_collision(osp.join(folderName, f))
        col += sum(len(x) for x in allfIdx)
        bcol += sum(int(np.sum(x)) for x in allbCol)
    return col, bcol


def save_collisions_seq(folderName, isSubFolder=True, outFolder=None):
    """Cut +-7-frame clips around every detected collision into a new dataset (utils.py:88-131)."""
    import scipy.io as sio
    names = sorted(f for f in os.listdir(folderName) if osp.isdir(osp.join(folderName, f)))
    if outFolder is None:
        outFolder = folderName.rstrip("/") + ("-collisions-only" if isSubFolder else "-collisions-only-nosub")
    count = 0
    for f in names:
        src = osp.join(folderName, f)
        allfIdx, _ = detect_collision(src)
        mat = sio.loadmat(osp.join(src, "data.mat"))
        pos, force = mat["position"], mat["force"]
        if isSubFolder:
            count = 0
        for c in [i for per_ball in allfIdx for i in per_ball]:
            st, en = max(0, c - 7), min(c + 7, pos.shape[1])
            dst = osp.join(outFolder, f, "sub-%04d" % count) if isSubFolder else osp.join(outFolder, "seq%06d" % count)
            os.makedirs(dst, exist_ok=True)
            count += 1
            for k, idx in enumerate(range(st, en)):
                shutil.copy(osp.join(src, "im%06d.jpg" % idx), osp.join(dst, "im%06d.jpg" % k))
            sio.savemat(osp.join(dst, "data.mat"), {"position": pos[:, st:en], "force": force[:, st:en]})
    return count
