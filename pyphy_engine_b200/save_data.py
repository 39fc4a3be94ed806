"""The reference's checked-in dataset driver (save_data.py:3-12): validation sets for 1, 2 and 3
balls with the seeds 3792 / 4185 / 1094, |F| = 8e4, 210-frame sequences, arena 700.

``save(rootPath=...)`` writes the same trees as the reference; every ``numSeq`` block is one GPU
batch (see ``dataio.DataSaver.save``)."""
from __future__ import annotations

from . import dataio as dio

NUM_BALLS = [1, 2, 3]
RAND_SEEDS = [3792, 4185, 1094]   # one per ball count, "val" prefix


def save(numSeq=100, rootPath="/data0/pulkitag/projPhysics/", **kwargs):
    for nb, seed in zip(NUM_BALLS, RAND_SEEDS):
        dio.save_rect_arena(numSeq=numSeq, numBalls=nb, svPrefix="val", mnForce=8e4, mxForce=8e4, mnSeqLen=210,
                            mxSeqLen=210, randSeed=seed, rootPath=rootPath, **kwargs)
