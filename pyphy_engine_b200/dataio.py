"""Frame-sequence data generation with the reference's ``dataio.py`` interface.

``DataSaver`` keeps the reference's constructor, ``save(numSeq)``, ``fetch(cropSz, procSz)``,
``generate_model()`` and output tree::

    <rootPath>/<expStr_>/seq%06d/{im%06d.jpg, data.mat, world.pkl}        (dataio.py:38-64, 88-129)

World sampling (walls, ball placement by rejection, initial forces) follows the reference draw by
draw on ``np.random.RandomState(randSeed)``, so a given seed produces the
same tables, balls and forces as the reference -- the golden fixtures pin that.  What differs is
the execution: the sequences of one ``save`` call are sampled first (sampling never looks at
simulation results, so the random stream is unchanged) and then stepped and drawn together as ONE
batch on the GPU -- every sequence is a world of a ``WorldBatch`` -- while host threads encode the
JPEGs.  ``fetch_batch`` returns the same data as arrays without touching the disk.

There is no CPU simulation or drawing in this module.
"""
from __future__ import annotations

import os
import pickle
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import geometry as gm
from . import primitives as pm

__all__ = ["DataSaver", "CustomWorld", "fix_theta_range", "save_rect_arena", "save_nonrect_arena_train", "save_nonrect_arena_val",
           "save_multishape_rect_arena"]


class SequenceSpec(object):
    """One sampled sequence: what ``generate_model`` produced plus its length."""

    def __init__(self, model, forces, ball_pos, cage_pts, seq_len):
        self.model, self.forces, self.ball_pos, self.cage_pts, self.seq_len = model, forces, ball_pos, cage_pts, seq_len


class DataSaver(object):
    def __init__(self, rootPath="/work5/pulkitag/projPhysics", numBalls=1, mnBallSz=15, mxBallSz=35, mnSeqLen=40,
                 mxSeqLen=100, mnForce=1e+3, mxForce=1e+6, wThick=30, isRect=True, wTheta=30, mxWLen=600,
                 mnWLen=200, arenaSz=667, oppForce=False, svPrefix=None, randSeed=None, verbose=0, dtype="f64",
                 device=0, makeDirs=True, **kwargs):
        """Reference arguments (dataio.py:21-36) plus: dtype "f64" (reference arithmetic) | "f32"
        (throughput), device (CUDA ordinal), makeDirs=False to skip creating the output tree."""
        self.expStr_ = "aSz%d_wLen%d-%d_nb%d_bSz%d-%d_f%.2e-%.2e_sLen%d-%d_wTh%d" % (
            arenaSz, mnWLen, mxWLen, numBalls, mnBallSz, mxBallSz, mnForce, mxForce, mnSeqLen, mxSeqLen, wThick)
        if svPrefix is not None:
            self.expStr_ = svPrefix + "-" + self.expStr_
        if not isRect:
            # the reference's suffix, malformed join included (dataio.py:44-50)
            if isinstance(wTheta, list):
                self.expStr_ += "_wTheta".join("%d-" % th for th in wTheta)[0:-1]
            else:
                self.expStr_ += "_wTheta%d" % wTheta
        if oppForce:
            self.expStr_ += "_oppFrc"
        self.dirName_ = os.path.join(rootPath, self.expStr_)
        if makeDirs and not os.path.exists(self.dirName_):
            os.makedirs(self.dirName_)
        self.seqDir_ = os.path.join(self.dirName_, "seq%06d")
        self.mnSeqLen_, self.mxSeqLen_ = mnSeqLen, mxSeqLen
        self.imFile_, self.dataFile_, self.worldFile_ = "im%06d.jpg", "data.mat", "world.pkl"
        self.numBalls_ = numBalls
        self.bmn_, self.bmx_ = mnBallSz, mxBallSz
        self.fmn_, self.fmx_ = mnForce, mxForce
        self.wlmx_, self.wlmn_ = mxWLen, mnWLen
        self.xSz_ = self.ySz_ = arenaSz
        self.wth_ = wThick
        self.isRect_ = isRect
        self.oppForce_ = oppForce
        self.verbose_ = verbose
        self.wTheta_ = wTheta if isinstance(wTheta, list) else [wTheta]
        self.rand_ = np.random.RandomState() if randSeed is None else np.random.RandomState(randSeed)
        self.dtype_ = dtype
        self.device_ = device
        self.seqLen_ = None

    # ------------------------------------------------------------------------------------------
    # sampling (host; identical random stream to the reference)
    # ------------------------------------------------------------------------------------------
    def _draw_seq_len(self):
        return int(self.mnSeqLen_ + self.rand_.rand() * (self.mxSeqLen_ - self.mnSeqLen_))   # dataio.py:91, 132

    def _generate_model(self, returnPos=False):
        self.world_ = pm.World(xSz=self.xSz_, ySz=self.ySz_, dtype=self.dtype_, device=self.device_)
        walls = self.add_walls()
        ballpos = self.add_balls()
        model = pm.Dynamics(self.world_)
        return (model, ballpos, self.pts, walls) if returnPos else model

    def generate_model(self):
        """-> (model, forces, ball positions, walls), dataio.py:213-217."""
        model, ballpos, _, walls = self._generate_model(returnPos=True)
        model, fs = self.apply_force(model)
        return model, fs, ballpos, walls

    def add_rectangular_walls(self, fColor=pm.Color(1.0, 0.0, 0.0)):
        """hlen, vlen, xleft, ytop: four draws (dataio.py:220-230)."""
        span = self.wlmx_ - self.wlmn_
        hlen = np.floor(self.wlmn_ + self.rand_.rand() * span)
        vlen = np.floor(self.wlmn_ + self.rand_.rand() * span)
        xleft = np.floor(self.rand_.rand() * (self.xSz_ - (hlen + self.wth_)))
        ytop = np.floor(self.rand_.rand() * (self.ySz_ - (vlen + self.wth_)))
        return self._create_walls(xleft, ytop, (0, 90, 180), (hlen - self.wth_, vlen, hlen - self.wth_), fColor=fColor)

    def sample_walls(self):
        """Rhombus table: angle, side length, position; resampled until it fits (dataio.py:243-274)."""
        while True:
            wtheta = self.wTheta_[self.rand_.permutation(len(self.wTheta_))[0]]
            rad = (wtheta * np.pi) / 180.0
            hlen = self.wlmn_ + self.rand_.rand() * (self.wlmx_ - self.wlmn_)
            if wtheta == 90:
                xlen = ylen = hlen
            else:
                xlen, ylen = hlen * np.cos(rad), hlen * np.sin(rad)
            xleftmin = self.wth_
            xleftmax = self.xSz_ - (2 * xlen + 2 * self.wth_)
            yleftmin = ylen + self.wth_
            yleftmax = self.ySz_ - self.wth_ if wtheta == 90 else self.ySz_ - (ylen + self.wth_)
            if xleftmin <= 0 or yleftmin <= 0 or xleftmax < xleftmin or yleftmax < yleftmin:
                continue
            xleft = xleftmin + np.floor(self.rand_.rand() * (xleftmax - xleftmin))
            yleft = yleftmin + np.floor(self.rand_.rand() * (yleftmax - yleftmin))
            return xleft, yleft, wtheta, hlen

    def _create_walls(self, xleft, yleft, thetas, wlens, fColor):
        pts = [gm.Point(xleft, yleft)]
        for th, wl in zip(thetas, wlens):
            pts.append(pts[-1] + (wl * gm.theta2dir(th)))
        walls = pm.create_cage(pts, wThick=self.wth_, fColor=fColor)
        self.pts = pts
        for w in walls:
            self.world_.add_object(w)
        self.lines_ = [gm.Line(pts[i], pts[(i + 1) % len(pts)]) for i in range(len(pts))]
        return walls

    def add_walls(self, fColor=pm.Color(1.0, 0.0, 0.0)):
        if self.isRect_:
            return self.add_rectangular_walls(fColor=fColor)
        xleft, yleft, wtheta, hlen = self.sample_walls()
        return self._create_walls(xleft, yleft, (-wtheta, wtheta, 180 - wtheta), (hlen, hlen, hlen), fColor=fColor)

    def find_point_within_lines(self, mindist):
        """Integer point whose signed distance from every cage line exceeds ``mindist`` (2 draws)."""
        x = int(np.round(self.pts[0].x() + self.rand_.rand() * (self.pts[2].x() - self.pts[0].x())))
        y = int(np.round(self.pts[1].y() + self.rand_.rand() * (self.pts[3].y() - self.pts[1].y())))
        pt = gm.Point(x, y)
        dist = [l.distance_to_point(pt) for l in self.lines_]
        return pt, all(d > mindist for d in dist), min(dist)

    def add_balls(self):
        """Rejection sampling of radii and centres (dataio.py:328-379)."""
        allr, allpos = [], []
        for i in range(self.numBalls_):
            while True:
                r = int(np.floor(self.bmn_ + self.rand_.rand() * (self.bmx_ - self.bmn_)))
                bdef = pm.BallDef(radius=r, fColor=pm.Color(0.5, 0.5, 0.5))
                tries = 0
                while True:
                    pt, ok, _ = self.find_point_within_lines(r + self.wth_ + 2)
                    tries += 1
                    if ok:
                        break
                    if tries >= 500:
                        raise RuntimeError("failed to find a point to place the ball (dataio.py:359-362)")
                pt = gm.Point(pt.x_asint(), pt.y_asint())
                if all(pt.distance(allpos[j]) > (allr[j] + r) for j in range(i)):
                    allr.append(r)
                    allpos.append(pt)
                    break
            self.world_.add_object(bdef, initPos=gm.Point(pt.x(), pt.y()))
        return allpos

    def apply_force(self, model):
        """Initial kick per ball (dataio.py:382-416); the direction reuses the magnitude's draw."""
        fs = []
        for i in range(self.numBalls_):
            name = "ball-%d" % i
            if self.oppForce_:
                p1 = self.world_.get_object_position(name)
                p2 = self.world_.get_object_position("ball-%d" % ((i + 1) % 2))
                ff = p2 - p1
                mag = self.fmn_ + np.floor(self.rand_.rand() * (self.fmx_ - self.fmn_))
                ff.make_unit_norm()
                ff.scale(mag)
                fx, fy = ff.x(), ff.y()
            else:
                rnd1, _rnd2 = self.rand_.rand(), self.rand_.rand()
                fmag = self.fmn_ + np.floor(rnd1 * (self.fmx_ - self.fmn_))
                ftheta = rnd1 * np.pi
                if self.rand_.rand() > 0.5:
                    ftheta = -ftheta
                fx, fy = fmag * np.cos(ftheta), fmag * np.sin(ftheta)
            f = gm.Point(fx, fy)
            model.apply_force(name, f, forceT=1.0)
            fs.append(f)
        return model, fs

    def sample_sequences(self, numSeq):
        """The random draws of ``numSeq`` consecutive ``save`` / ``fetch`` iterations."""
        specs = []
        for _ in range(numSeq):
            self.seqLen_ = self._draw_seq_len()
            model, fs, ballpos, _ = self.generate_model()
            specs.append(SequenceSpec(model, fs, ballpos, self.pts, self.seqLen_))
        return specs

    # ------------------------------------------------------------------------------------------
    # batched execution
    # ------------------------------------------------------------------------------------------
    def _batch_from_specs(self, specs, want_frames):
        from .engine import WorldBatch
        n, nB = len(specs), self.numBalls_
        nW = len(specs[0].model.world_.static_)
        wall = np.zeros((n, nW, 4, 2))
        pos = np.zeros((n, nB, 2))
        vel = np.zeros((n, nB, 2))
        rad = np.zeros((n, nB))
        mass = np.zeros((n, nB))
        for s, sp in enumerate(specs):
            world = sp.model.world_
            for i, w in enumerate(world.static_.values()):
                for k, v in enumerate(w.bbox_.vert_):
                    wall[s, i, k] = (v.x(), v.y())
            for j, b in enumerate(world.dynamic_.values()):
                pos[s, j] = (b.pos_.x(), b.pos_.y())
                vel[s, j] = (b.vel_.x(), b.vel_.y())
                rad[s, j], mass[s, j] = b.get_radius(), b.get_mass()
        canvas = (int(self.xSz_), int(self.ySz_)) if want_frames else (0, 0)
        wb = WorldBatch(n, nB, nW, dtype=self.dtype_, canvas=canvas, device=self.device_)
        wb.set_walls(wall)
        wb.set_balls(pos, vel, rad, mass)
        if want_frames:
            b0 = next(iter(specs[0].model.world_.dynamic_.values()))
            wb.set_styles([(int(b0.sThick_), b0.fColor_.rgba(), b0.sColor_.rgba())] * nB,
                          [(0, (1.0, 0.0, 0.0, 1.0))] * nW, int(rad.min()), int(rad.max()))
        return wb

    def run_sequences(self, specs, want_frames=True, frame_sink=None, max_stage_bytes=1 << 30):
        """Step (and draw) all sequences as one batch.

        Returns ``position`` -- list of (2*numBalls, seqLen) float32 arrays, ``position[2j:2j+2, i]`` =
        ball j after step i+1 (dataio.py:123-128) -- and, when ``frame_sink`` is None, the frames as
        a list (per sequence) of (seqLen, H, W, 4) uint8 arrays.  ``frame_sink(seq, step, image)`` is
        called for every frame otherwise (possibly from worker threads).
        """
        import torch
        H, W = int(self.ySz_), int(self.xSz_)
        # a batch is bounded by one step of frames fitting one of the two staging buffers: larger jobs run as several batches
        group = max(1, (max_stage_bytes // 2) // (H * W * 4)) if want_frames else 1 << 20
        if len(specs) > group:
            position, frames_out = [], ([] if (want_frames and frame_sink is None) else None)
            for g0 in range(0, len(specs), group):
                sink = None if frame_sink is None else (lambda s, i, im, g0=g0: frame_sink(g0 + s, i, im))
                p, f = self.run_sequences(specs[g0:g0 + group], want_frames, sink, max_stage_bytes)
                position += p
                if frames_out is not None:
                    frames_out += f
            return position, frames_out
        n, nB = len(specs), self.numBalls_
        lens = np.array([sp.seq_len for sp in specs])
        total = int(lens.max()) if n else 0
        wb = self._batch_from_specs(specs, want_frames)
        position = [np.zeros((2 * nB, int(l)), np.float32) for l in lens]
        frames_out = [np.empty((int(l), H, W, 4), np.uint8) for l in lens] if (want_frames and frame_sink is None) else None
        per_step = n * nB * 2 * 4 + (n * H * W * 4 if want_frames else 0)
        # two page-locked staging sets: while the host scatters chunk i into the per-sequence arrays (several
        # threads: first-touch page faults of the destination dominate that copy), the GPU already steps, draws
        # and downloads chunk i+1 into the other set (ppb_simulate_host_async / ppb_host_sync)
        k = int(max(1, min(total, (max_stage_bytes // 2) // max(per_step, 1))))
        n_sets = 1 if total <= k else 2           # a job of one chunk has nothing to overlap with
        traj_host = [torch.empty((k, n, nB, 2), dtype=torch.float32).pin_memory() for _ in range(n_sets)]
        frames_host = [torch.empty((k, n, H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(n_sets)] if want_frames else None
        pool = ThreadPoolExecutor(max_workers=min(8, max(1, n)))

        def scatter(buf, done, cur):
            tr = traj_host[buf][:cur].numpy()

            def one(s):
                m = int(min(cur, lens[s] - done))
                if m <= 0:
                    return
                position[s][:, done:done + m] = tr[:m, s].reshape(m, 2 * nB).T
                if want_frames:
                    fr = frames_host[buf][:m, s].numpy()
                    if frame_sink is None:
                        frames_out[s][done:done + m] = fr
                    else:
                        for i in range(m):
                            frame_sink(s, done + i, fr[i].copy())
            list(pool.map(one, range(n)))

        try:
            done, it, pending = 0, 0, None
            while done < total:
                cur = min(k, total - done)
                buf = it & (n_sets - 1)
                wb.simulate_host(cur, traj_host[buf][:cur], frames_host[buf][:cur] if want_frames else None, 1, wait=False)
                if pending is not None:
                    wb.host_sync(keep_in_flight=1)          # the previous chunk is complete on the host
                    scatter(*pending)
                pending = (buf, done, cur)
                done += cur
                it += 1
            if pending is not None:
                wb.host_sync()
                scatter(*pending)
            status, at = wb.get_status()
            bad = [(s, int(status[s]), int(at[s])) for s in range(n) if status[s] != 0 and at[s] < lens[s]]
            if bad:
                from ._capi import STATUS_NAMES
                s, st, step = bad[0]
                raise AssertionError("sequence %d: the reference raises in step %d (%s); %d sequence(s) affected"
                                     % (s, step, STATUS_NAMES.get(st, st), len(bad)))
        finally:
            pool.shutdown(wait=True)
            wb.close()
        return position, frames_out

    # ------------------------------------------------------------------------------------------
    # reference entry points
    # ------------------------------------------------------------------------------------------
    def save(self, numSeq=10, workers=8):
        """Writes ``numSeq`` sequences (dataio.py:88-129)."""
        import scipy.io as sio
        from PIL import Image
        specs = self.sample_sequences(numSeq)
        dirs = []
        for i, sp in enumerate(specs):
            d = self.seqDir_ % i
            os.makedirs(d, exist_ok=True)
            dirs.append(d)
            with open(os.path.join(d, self.worldFile_), "wb") as fh:
                pickle.dump({"force": sp.forces, "ballPos": sp.ball_pos, "walls": sp.cage_pts}, fh, 2)
        pool = ThreadPoolExecutor(max_workers=workers)
        jobs = []

        def write_jpeg(path, im):
            Image.fromarray(im[:, :, :3]).save(path)     # scm.imsave(name.jpg, im): PIL defaults

        def sink(seq, step, im):
            jobs.append(pool.submit(write_jpeg, os.path.join(dirs[seq], self.imFile_ % step), im))

        try:
            position, _ = self.run_sequences(specs, want_frames=True, frame_sink=sink)
        finally:
            for j in jobs:
                j.result()
            pool.shutdown()
        for i, sp in enumerate(specs):
            force = np.zeros((2 * self.numBalls_, sp.seq_len), np.float32)
            for b, fb in enumerate(sp.forces):
                force[2 * b, 0], force[2 * b + 1, 0] = fb.x(), fb.y()
            sio.savemat(os.path.join(dirs[i], self.dataFile_), {"force": force, "position": position[i]})

    def save_sequence(self, dataFile, imFile, worldFile):
        """One sequence into explicit file names (dataio.py:101-129); ``self.seqLen_`` must be set."""
        import scipy.io as sio
        from PIL import Image
        model, f, ballPos, _ = self.generate_model()
        sp = SequenceSpec(model, f, ballPos, self.pts, self.seqLen_)
        with open(worldFile, "wb") as fh:
            pickle.dump({"force": f, "ballPos": ballPos, "walls": self.pts}, fh, 2)
        position, frames = self.run_sequences([sp], want_frames=True)
        for i in range(sp.seq_len):
            Image.fromarray(frames[0][i][:, :, :3]).save(imFile % i)
        force = np.zeros((2 * self.numBalls_, sp.seq_len), np.float32)
        for b, fb in enumerate(f):
            force[2 * b, 0], force[2 * b + 1, 0] = fb.x(), fb.y()
        sio.savemat(dataFile, {"force": force, "position": position[0]})

    def fetch(self, cropSz=None, procSz=None):
        """One sequence in memory (dataio.py:131-196): the list of frames, or with ``cropSz`` the
        ball-centred crops ``(imBalls, force, velocity, position)``.  Stepping, drawing, cropping and the
        optional ``procSz`` resize (scipy.misc.imresize in the reference, i.e. PIL bilinear) run on the device
        (``ppb_step_async`` + ``ppb_crop_async`` + ``ppb_resize_crops_async``)."""
        self.seqLen_ = self._draw_seq_len()
        model, f, ballPos, _ = self.generate_model()
        sp = SequenceSpec(model, f, ballPos, self.pts, self.seqLen_)
        if cropSz is None:
            _, frames = self.run_sequences([sp], want_frames=True)
            return [frames[0][i] for i in range(sp.seq_len)]
        return self._fetch_crops(sp, int(cropSz), procSz)

    def _fetch_crops(self, sp, cropSz, procSz):
        nB, L = self.numBalls_, sp.seq_len
        wb = self._batch_from_specs([sp], True)
        try:
            traj, frames = wb.step(L, record_traj=True, render_every=1)
            status, at = wb.get_status()
            if status[0] != 0 and at[0] < L:
                from ._capi import STATUS_NAMES
                raise AssertionError("the reference raises in step %d (%s)" % (int(at[0]), STATUS_NAMES.get(int(status[0]))))
            crops, cpos = wb.crop(frames, traj, cropSz)
            vel = None
            if procSz is not None:                                  # dataio.py:183-189, PIL bilinear on the device
                crops, cpos, vel = wb.resize_crops(crops, int(procSz), pos=cpos, traj=traj, want_vel=True)
                vel = vel[:, 0].cpu().numpy()                       # (L, nB, 2); row i-1 = frame i - frame i-1, scaled
            crops = crops[:, 0].cpu().numpy()                       # (L, nB, c, c, 3)
            cpos = cpos[:, 0].cpu().numpy()                         # (L, nB, 2) in crop coordinates
        finally:
            wb.close()
        force = np.zeros((2 * nB, L), np.float32)
        for b, fb in enumerate(sp.forces):
            force[2 * b, 0], force[2 * b + 1, 0] = fb.x(), fb.y()
        position = cpos.reshape(L, 2 * nB).T.copy()
        velocity = np.zeros((2 * nB, L), np.float32)
        imBalls = [[crops[i, j] for i in range(L)] for j in range(nB)]
        if vel is not None:
            velocity[:, :L - 1] = vel[:L - 1].reshape(L - 1, 2 * nB).T
        return imBalls, force, velocity, position

    def fetch_batch(self, numSeq, want_frames=True):
        """``numSeq`` fetches as one GPU batch: (specs, position list, frames list)."""
        specs = self.sample_sequences(numSeq)
        position, frames = self.run_sequences(specs, want_frames=want_frames)
        return specs, position, frames


class CustomWorld(object):
    """A world with a fixed layout -- given ball centres and kicks in a cage of side ``wlen`` -- and the
    per-ball "glimpse" of it (dataio.py:457-570).  ``get_glimpse`` steps the world, draws it and cuts the
    ball-centred crops on the device (``ppb_crop_async`` mode 1: the offsets are truncated with ``int()``
    as in dataio.py:510-514, unlike ``DataSaver.fetch`` which rounds them)."""

    def __init__(self, numballs=1, ballsz=25, wthick=30, isrect=True, wtheta=30, wlen=300, arenasz=667, verbose=0,
                 ballLocs=None, forces=None, dtype="f64", device=0, **kwargs):
        self.numballs_ = numballs
        self.bsz_ = ballsz
        self.bloc_ = [gm.Point.from_self(p) for p in (ballLocs if ballLocs is not None else [gm.Point(120, 120)])]
        self.forces_ = [gm.Point.from_self(f) for f in (forces if forces is not None else [gm.Point(5e+4, 5e+4)])]
        self.wthick_ = wthick
        self.wtheta_ = wtheta
        self.isrect_ = isrect
        self.wlen_ = wlen
        self.asz_ = arenasz
        self.verbose_ = verbose
        self.xSz_ = self.ySz_ = arenasz
        self.dtype_, self.device_ = dtype, device

    def _generate_model(self):
        self.world_ = pm.World(xSz=self.asz_, ySz=self.asz_, dtype=self.dtype_, device=self.device_)
        self.walls_ = self.add_walls()
        self.add_balls()
        self.model_ = pm.Dynamics(self.world_)

    def generate_model(self):
        self._generate_model()
        self.apply_force()

    def apply_force(self):
        for b in range(self.numballs_):
            self.model_.apply_force("ball-%d" % b, self.forces_[b], forceT=1.0)

    def add_walls(self, xleft=20, yleft=20, hLen=None, vLen=None, fColor=pm.Color(1.0, 0.0, 0.0)):
        hLen = self.wlen_ if hLen is None else hLen
        vLen = self.wlen_ if vLen is None else vLen
        wt = self.wtheta_
        th = list(wt) if isinstance(wt, list) else [-wt, wt, 180 - wt]
        th = [fix_theta_range(t) for t in th]
        return self._create_walls(xleft, yleft, th, (vLen, hLen, vLen), fColor=fColor)

    def _create_walls(self, xleft, yleft, thetas, wlens, fColor):
        pts = [gm.Point(xleft, yleft)]
        for t, wl in zip(thetas, wlens):
            pts.append(pts[-1] + (wl * gm.theta2dir(t)))
        walls = pm.create_cage(pts, wThick=self.wthick_, fColor=fColor)
        self.pts = pts
        for w in walls:
            self.world_.add_object(w)
        self.lines_ = [gm.Line(pts[i], pts[(i + 1) % len(pts)]) for i in range(len(pts))]
        return walls

    def add_balls(self):
        for i in range(self.numballs_):
            self.world_.add_object(pm.BallDef(radius=self.bsz_, fColor=pm.Color(0.5, 0.5, 0.5)), initPos=self.bloc_[i])

    def get_glimpse(self, ballNum=None, cropSz=128):
        """step() + generate_image() + one white-padded crop per ball: (imBalls, position, velocity)."""
        import torch
        self.model_.step()
        nb = self.numballs_
        position, velocity = np.zeros((2 * nb,)), np.zeros((2 * nb,))
        for b in range(nb):
            ball = self.model_.get_object("ball-%d" % b)
            p, v = ball.get_position(), ball.get_velocity()
            position[2 * b], position[2 * b + 1] = p.x(), p.y()
            velocity[2 * b], velocity[2 * b + 1] = v.x(), v.y()
        imBalls = []
        if cropSz is not None:
            dev = self.model_._dev()
            dev.ensure_styles(self.world_)
            frame = dev.batch.render()                                        # (1, H, W, 4) on the device
            tdt = torch.float64 if self.dtype_ in ("f64", "float64", "fp64") else torch.float32
            centres = torch.from_numpy(position.reshape(1, 1, nb, 2)).to(frame.device, tdt).contiguous()
            crops, _ = dev.batch.crop(frame[None], centres, int(cropSz), mode=1)
            crops = crops[0, 0].cpu().numpy()
            imBalls = [crops[b] for b in range(nb)]
        return imBalls, position, velocity


def fix_theta_range(theta):
    """Angle in degrees folded into (-180, 180] (dataio.py:573-577)."""
    theta = np.mod(theta, 360)
    if theta > 180:
        theta = -(360 - theta)
    return theta


# ----------------------------------------------------------------------------------------------
# the reference's preset drivers (dataio.py:419-453; the lower-case ``datasaver`` typos fixed)
# ----------------------------------------------------------------------------------------------
def save_rect_arena(numSeq=10000, oppForce=False, numBalls=1, svPrefix=None, mnForce=3e+4, mxForce=8e+4, mnWLen=300,
                    mxWLen=550, arenaSz=700, mnSeqLen=10, mxSeqLen=200, randSeed=None,
                    rootPath="/data0/pulkitag/projPhysics/", **kwargs):
    sv = DataSaver(wThick=30, isRect=True, mnForce=mnForce, mxForce=mxForce, mnWLen=mnWLen, mxWLen=mxWLen,
                   mnSeqLen=mnSeqLen, mxSeqLen=mxSeqLen, numBalls=numBalls, mnBallSz=25, mxBallSz=25, arenaSz=arenaSz,
                   oppForce=oppForce, svPrefix=svPrefix, rootPath=rootPath, randSeed=randSeed, **kwargs)
    sv.save(numSeq=numSeq)
    return sv


def save_nonrect_arena_val(numSeq=100, rootPath="/work5/pulkitag/projPhysics", **kwargs):
    sv = DataSaver(rootPath=rootPath, wThick=20, isRect=False, mxForce=1e+5, mnSeqLen=10, mxSeqLen=100,
                   wTheta=[23, 38, 45, 53], **kwargs)
    sv.save(numSeq=numSeq)
    return sv


def save_nonrect_arena_train(numSeq=10000, oppForce=False, numBalls=1, svPrefix=None, mnWLen=500, mxWLen=800,
                             arenaSz=1600, mnForce=3e+4, mxForce=8e+4, rootPath="/data0/pulkitag/projPhysics/", **kwargs):
    sv = DataSaver(rootPath=rootPath, wThick=30, isRect=False, mnForce=mnForce, mxForce=mxForce, mnWLen=mnWLen,
                   mxWLen=mxWLen, numBalls=numBalls, mnSeqLen=10, mxSeqLen=200, mnBallSz=25, mxBallSz=25,
                   wTheta=[30, 60], arenaSz=arenaSz, svPrefix=svPrefix, oppForce=oppForce, **kwargs)
    sv.save(numSeq=numSeq)
    return sv


def save_multishape_rect_arena(numSeq=1000, numBalls=1, oppForce=False, rootPath="/data1/pulkitag/projPhysics/", **kwargs):
    sv = DataSaver(rootPath=rootPath, numBalls=numBalls, wThick=20, isRect=True, mnForce=1e+4, mxForce=1e+5,
                   mnWLen=400, mxWLen=400, mnSeqLen=40, mxSeqLen=40, mnBallSz=25, mxBallSz=25, oppForce=oppForce,
                   **kwargs)
    sv.save(numSeq=numSeq)
    return sv
