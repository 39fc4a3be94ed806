"""Ball / wall objects, ``World`` and ``Dynamics`` with the reference's ``primitives.py`` interface.

Drop-in for the reference module on the simulate-and-render path::

    import pyphy_engine_b200.geometry as gm
    import pyphy_engine_b200.primitives as pm

    world = pm.World(xSz=64, ySz=64)
    for wall in pm.create_cage([gm.Point(2, 2), gm.Point(62, 2), gm.Point(62, 62), gm.Point(2, 62)], wThick=3):
        world.add_object(wall)
    world.add_object(pm.BallDef(radius=4, sThick=1, fColor=pm.Color(.5, .5, .5)), initPos=gm.Point(20, 30))
    model = pm.Dynamics(world)
    model.set_object_velocity('ball-0', gm.Point(60, 25))
    model.step()
    im = model.generate_image()          # (ySz, xSz, 4) uint8

The objects are plain Python descriptions of a world; ``Dynamics.step`` and
``World.generate_image`` run on the GPU through the C ABI (``include/pyphy_b200.h``): a
``World`` is row 0 of a one-world ``WorldBatch``.  Batched use (thousands of worlds
per call) goes through ``pyphy_engine_b200.engine.WorldBatch`` / ``dataio.DataSaver``.
There is no CPU stepping or drawing code in this module.

Reference quirks that are kept on purpose: ``apply_force`` adds
``0.5*T^2*F/m`` to the velocity and does not refresh collision caches (primitives.py:580-594);
mass ignores ``BallDef.density`` (:394-400); ``GenericWall`` is always drawn red because its
sprite is built without the ``fColor`` argument (:232-234); names are ``wall-%d`` / ``ball-%d``
in insertion order (:905-933).
"""
from __future__ import annotations

import collections as co
import copy
import pickle
import warnings
import weakref

import numpy as np

from . import geometry as gm

__all__ = ["Color", "CairoData", "get_box_coords", "create_cage", "find_top_left", "WallDef", "GenericWall", "Wall",
           "BallDef", "Ball", "World", "Dynamics", "DynamicsHorizon", "get_ball_im", "get_rectangle_im"]


class Color(object):
    def __init__(self, r, g, b, a=1.0):
        self.r, self.g, self.b, self.a = r, g, b, a

    def rgba(self):
        return (float(self.r), float(self.g), float(self.b), float(self.a))


class CairoData(object):
    """(context, image) pair kept for interface compatibility; ``cr`` is always None here."""

    def __init__(self, cr, im):
        self.cr = cr
        self.im = im


_RED = (1.0, 0.0, 0.0, 1.0)


# ------------------------------------------------------------------------------------------
# cage construction (primitives.py:177-209)
# ------------------------------------------------------------------------------------------
def get_box_coords(stPoint, enPoint, wThick=30):
    """The segment st->en thickened by ``wThick`` on the side opposite to the line's
    outward normal: [st, st + L d, that - T n, that - L d] (primitives.py:177-195)."""
    line = gm.Line(stPoint, enPoint)
    length = (enPoint - stPoint).mag()
    n = line.get_normal()
    d = line.get_direction()
    p1 = stPoint
    p2 = p1 + length * d
    p3 = p2 - (wThick * n)
    p4 = p3 - (length * d)
    return [p1, p2, p3, p4]


def create_cage(pts, wThick=30, fColor=Color(1.0, 0.0, 0.0)):
    """One GenericWall per polygon side pts[i] -> pts[i+1] (primitives.py:199-209)."""
    n = len(pts)
    return [GenericWall(pts[i], pts[(i + 1) % n], wThick=wThick, fColor=fColor) for i in range(n)]


def find_top_left(pts):
    return gm.Point(min([np.inf] + [p.x() for p in pts]), min([np.inf] + [p.y() for p in pts]))


# ------------------------------------------------------------------------------------------
# objects
# ------------------------------------------------------------------------------------------
class WallDef(object):
    """Axis-aligned rectangular wall: physical/visual definition only (primitives.py:213-218)."""

    def __init__(self, sz=gm.Point(4, 100), fColor=Color(1.0, 0.0, 0.0), name=None):
        self.sz = sz
        self.fColor = fColor
        self.name = name


class GenericWall(object):
    """A thickened segment: 4-vertex polygon with 4 collision edges (primitives.py:221-260)."""

    kind_ = 0   # ppb_wall_style.kind

    def __init__(self, stPoint, enPoint, fColor=Color(1.0, 0.0, 0.0), name=None, wThick=4):
        self.pts_ = get_box_coords(stPoint, enPoint, wThick=wThick)
        self.th_ = wThick
        self.pos_ = find_top_left(self.pts_)
        self.fColor_ = fColor          # kept, but the reference never draws with it (see module docstring)
        self.name_ = name
        self.bbox_ = gm.Bbox.from_list(self.pts_)
        ext = enPoint - stPoint
        self.block_dir_ = ext

    def get_lines(self):
        return self.bbox_.get_lines()

    def draw_rgba(self):
        return _RED


class Wall(object):
    """Axis-aligned rectangle at ``initPos`` (upper-left corner) of size ``sz`` (primitives.py:266-357)."""

    kind_ = 1

    def __init__(self, initPos=gm.Point(0, 0), sz=gm.Point(4, 100), fColor=Color(1.0, 0.0, 0.0), name=None):
        self.sz_ = sz
        self.pos_ = initPos
        self.fColor_ = fColor
        self.name_ = name
        self.make_coordinates()

    @classmethod
    def from_def(cls, wallDef, name, initPos):
        return cls(sz=wallDef.sz, fColor=wallDef.fColor, initPos=initPos, name=name)

    def make_coordinates(self):
        p, s = self.pos_, self.sz_
        self.lTop_ = p
        self.lBot_ = p + gm.Point(0, s.y())
        self.rTop_ = p + gm.Point(s.x(), 0)
        self.rBot_ = p + gm.Point(s.x(), s.y())
        self.l1_ = gm.Line(self.lTop_, self.lBot_)
        self.l2_ = gm.Line(self.lBot_, self.rBot_)
        self.l3_ = gm.Line(self.rBot_, self.rTop_)
        self.l4_ = gm.Line(self.rTop_, self.lTop_)
        self.bbox_ = gm.Bbox(self.lTop_, self.lBot_, self.rBot_, self.rTop_)

    def get_lines(self):
        return self.bbox_.get_lines()

    def get_collision_normal(self, pt):
        """Normal of the face ``pt`` looks at (primitives.py:320-345)."""
        loc = [l.get_point_location(pt) for l in (self.l1_, self.l2_, self.l3_, self.l4_)]
        if (loc[0] == -1 and loc[2] == 1) or loc[0] == 0:
            return self.l1_.get_normal()
        if (loc[0] == 1 and loc[2] == -1) or loc[2] == 0:
            return self.l3_.get_normal()
        if (loc[1] == 1 and loc[3] == -1) or loc[3] == 0:
            return self.l4_.get_normal()
        if (loc[1] == -1 and loc[3] == 1) or loc[1] == 0:
            return self.l2_.get_normal()
        return None

    def check_collision(self, headingDir):
        return self.bbox_.get_intersection_with_line_ray(headingDir)

    def name(self):
        return self.name_

    def draw_rgba(self):
        return self.fColor_.rgba()


class BallDef(object):
    def __init__(self, radius=20, sThick=2, sColor=Color(0.0, 0.0, 0.0), fColor=Color(1.0, 0.0, 0.0), name=None,
                 density=1.0):
        self.radius = radius
        self.sThick = sThick
        self.sColor = sColor
        self.fColor = fColor
        self.name = name
        self.density = density


class Ball(object):
    """primitives.py:375-501.  mass = 4/3 pi (r/10)^3 density."""

    def __init__(self, radius=20, sThick=2, sColor=Color(0.0, 0.0, 0.0), fColor=Color(1.0, 0.0, 0.0), name=None,
                 initPos=gm.Point(0, 0), initVel=gm.Point(0, 0), density=1.0):
        self.radius_ = radius
        self.sThick_ = sThick
        self.sColor_ = sColor
        self.fColor_ = fColor
        self.name_ = name
        self.pos_ = initPos
        self.tCol_ = 0
        self.vel_ = initVel
        self.futureVel_ = gm.Point(0, 0)
        self.density_ = density
        self.mass_ = (4 / 3.0) * np.pi * np.power(self.radius_ / 10.0, 3) * density
        self.xSz_ = self.ySz_ = 2 * (radius + sThick)
        self.xOff_ = self.yOff_ = np.floor(self.xSz_ / 2)

    @classmethod
    def from_def(cls, ballDef, name, initPos, initVel=gm.Point(0, 0)):
        # density is not forwarded (primitives.py:394-400): every ball has density 1
        self = cls(radius=ballDef.radius, sThick=ballDef.sThick, sColor=ballDef.sColor, fColor=ballDef.fColor)
        self.name_ = name
        self.pos_ = initPos
        self.vel_ = initVel
        return self

    @classmethod
    def from_self(cls, other):
        self = cls()
        for k, v in vars(other).items():
            setattr(self, k, v)
        return self

    def set_name(self, name):
        self.name_ = name

    def name(self):
        return self.name_

    def get_position(self):
        return gm.Point.from_self(self.pos_)

    def get_mutable_position(self):
        return self.pos_

    def get_velocity(self):
        return gm.Point.from_self(self.vel_)

    def get_mutable_velocity(self):
        return self.vel_

    def set_position(self, pos):
        self.pos_ = gm.Point.from_self(pos)

    def set_velocity(self, vel):
        self.vel_ = gm.Point.from_self(vel)

    def get_mass(self):
        return self.mass_

    def set_mass(self, mass):
        self.mass_ = mass

    def get_radius(self):
        return self.radius_

    def set_after_collision_velocity(self, vel):
        self.futureVel_ = gm.Point.from_self(vel)

    def set_future_to_current_velocity(self):
        self.vel_ = gm.Point.from_self(self.futureVel_)

    def get_heading_direction_line(self):
        return gm.Line(self.pos_, self.pos_ + self.vel_)

    def get_point_of_contact(self, pt):
        head = self.get_heading_direction_line()
        assert head.get_point_location(pt) == 0, "The pt of collision and heading direction donot match"
        return head.get_point_along_line(self.pos_, self.radius_)


# ------------------------------------------------------------------------------------------
# device mirror of one World
# ------------------------------------------------------------------------------------------
class _DeviceWorld(object):
    """Row 0 of a one-world WorldBatch that mirrors ``world``'s objects.

    Rebuilt whenever the set of objects (or dt / dtype / canvas) changes -- which, as in the
    reference's ``Dynamics.add_object`` / ``del_object``, resets every collision cache and queues
    every ball.  Ball positions / velocities edited on the host between steps are pushed with
    ``update_balls`` (caches untouched, primitives.py:462-466).
    """

    def __init__(self, world, dt, g, dtype, device, event_capacity):
        from .engine import WorldBatch
        self.sig = self.signature(world, dt, dtype, device)
        walls, balls, order = self._split(world)
        self.wall_names = [n for n, _ in walls]
        self.ball_names = [n for n, _ in balls]
        nW, nB = len(walls), len(balls)
        if nB < 1:
            raise Exception("a world needs at least one ball before it can be stepped or drawn")
        self.batch = WorldBatch(1, nB, nW, dtype=dtype, canvas=(int(world.xSz_), int(world.ySz_)), dt=dt,
                                g=(g.x(), g.y()), order=order, device=device, event_capacity=event_capacity)
        self.g = (g.x(), g.y())
        wv = np.zeros((1, nW, 4, 2), np.float64)
        for i, (_, w) in enumerate(walls):
            for k, v in enumerate(w.bbox_.vert_):
                wv[0, i, k] = (v.x(), v.y())
        self.batch.set_walls(wv)
        pos, vel = self._host_state(world)
        rad = np.array([[b.get_radius() for _, b in balls]], np.float64)
        mass = np.array([[b.get_mass() for _, b in balls]], np.float64)
        self.batch.set_balls(pos, vel, rad, mass)
        self.last_pos, self.last_vel = pos, vel
        self.styles_ready = False
        self.steps = 0

    @staticmethod
    def _split(world):
        walls, balls, names = [], [], []
        for name, obj in world.objects_.items():
            names.append(name)
            (balls if isinstance(obj, Ball) else walls).append((name, obj))
        wi = {n: i for i, (n, _) in enumerate(walls)}
        bi = {n: i for i, (n, _) in enumerate(balls)}
        order = [wi[n] if n in wi else len(walls) + bi[n] for n in names]
        return walls, balls, order

    @staticmethod
    def signature(world, dt, dtype, device):
        items = []
        for name, obj in world.objects_.items():
            if isinstance(obj, Ball):
                items.append((name, "b", float(obj.get_radius()), float(obj.get_mass()), obj.sThick_,
                              obj.fColor_.rgba(), obj.sColor_.rgba()))
            else:
                items.append((name, "w", tuple((v.x(), v.y()) for v in obj.bbox_.vert_), obj.kind_, obj.draw_rgba()))
        return (tuple(items), float(dt), dtype, int(device), int(world.xSz_), int(world.ySz_))

    def _host_state(self, world):
        balls = [world.objects_[n] for n in self.ball_names] if hasattr(self, "ball_names") else \
            [o for o in world.objects_.values() if isinstance(o, Ball)]
        pos = np.array([[[b.pos_.x(), b.pos_.y()] for b in balls]], np.float64)
        vel = np.array([[[b.vel_.x(), b.vel_.y()] for b in balls]], np.float64)
        return pos, vel

    def push(self, world):
        pos, vel = self._host_state(world)
        dp = not np.array_equal(pos, self.last_pos)
        dv = not np.array_equal(vel, self.last_vel)
        if dp or dv:
            self.batch.update_balls(pos if dp else None, vel if dv else None)
            self.last_pos, self.last_vel = pos, vel

    def pull(self, world):
        pos, vel = self.batch.get_balls()
        for j, n in enumerate(self.ball_names):
            b = world.objects_[n]
            b.pos_ = gm.Point(pos[0, j, 0], pos[0, j, 1])
            b.vel_ = gm.Point(vel[0, j, 0], vel[0, j, 1])
        self.last_pos, self.last_vel = pos, vel

    def set_g(self, g):
        if (g.x(), g.y()) != self.g:
            self.batch.set_gravity(g.x(), g.y())
            self.g = (g.x(), g.y())

    def ensure_styles(self, world):
        if self.styles_ready:
            return
        balls = [world.objects_[n] for n in self.ball_names]
        walls = [world.objects_[n] for n in self.wall_names]
        radii = []
        for b in balls:
            r = b.get_radius()
            if int(r) != r or int(b.sThick_) != b.sThick_:
                raise ValueError("drawing needs integer ball radius and stroke (sprite side = 2*(radius+sThick), "
                                 "primitives.py:39-40); got radius=%r sThick=%r" % (r, b.sThick_))
            radii.append(int(r))
        self.batch.set_styles([(int(b.sThick_), b.fColor_.rgba(), b.sColor_.rgba()) for b in balls],
                              [(w.kind_, w.draw_rgba()) for w in walls], min(radii), max(radii))
        self.styles_ready = True

    def close(self):
        self.batch.close()


def _status_error(status, step):
    from ._capi import STATUS_NAMES
    return AssertionError("the reference would have raised in step %d: %s" % (step, STATUS_NAMES.get(int(status), status)))


# ------------------------------------------------------------------------------------------
# World
# ------------------------------------------------------------------------------------------
class World(object):
    """Ordered registries of static / dynamic objects + frame generation (primitives.py:856-1008).

    ``dtype`` ("f64" reference arithmetic, "f32" throughput) and ``device`` are extensions with
    reference-compatible defaults.
    """

    def __init__(self, xSz=200, ySz=200, deltaT=0.1, dtype="f64", device=0):
        self.static_ = co.OrderedDict()
        self.dynamic_ = co.OrderedDict()
        self.objects_ = co.OrderedDict()
        self.auxObjects_ = co.OrderedDict()
        self.count_ = co.OrderedDict()
        self.init_count()
        self.xSz_ = xSz
        self.ySz_ = ySz
        self.dtype_ = dtype
        self.device_ = device
        self._dev = None
        self._dev_params = (0.01, gm.Point(0, 0), 0)   # dt, g, event capacity used when only drawing

    # -- registries ----------------------------------------------------------------------------
    def init_count(self):
        self.count_["wall"] = 0
        self.count_["ball"] = 0

    def get_dynamic_object_names(self):
        return list(self.dynamic_.keys())

    def get_object_names(self):
        return list(self.objects_.keys())

    def get_object(self, name):
        assert name in self.objects_, "Object: %s not found" % name
        return self.objects_[name]

    def get_object_name_type(self, objDef, initPos=None):
        if isinstance(objDef, WallDef):
            name = "wall-%d" % self.count_["wall"]
            obj, kind = Wall.from_def(objDef, name, initPos), "static"
            self.count_["wall"] += 1
        elif isinstance(objDef, GenericWall):
            name = "wall-%d" % self.count_["wall"]
            obj, kind = objDef, "static"
            self.count_["wall"] += 1
        elif isinstance(objDef, BallDef):
            name = "ball-%d" % self.count_["ball"]
            obj, kind = Ball.from_def(objDef, name, initPos), "dynamic"
            self.count_["ball"] += 1
        elif isinstance(objDef, Ball):
            name = "ball-%d" % self.count_["ball"]
            obj, kind = Ball.from_self(objDef), "dynamic"
            if initPos is None:
                initPos = obj.get_position()
            obj.set_position(initPos)
            obj.set_name(name)
            self.count_["ball"] += 1
        else:
            raise Exception("Unrecognized object type")
        return obj, name, kind

    def add_object(self, objDef, initPos=gm.Point(0, 0)):
        obj, name, kind = self.get_object_name_type(objDef, initPos)
        if name in self.objects_:
            raise Exception("Object with name %s already exists" % name)
        self.objects_[name] = obj
        (self.static_ if kind == "static" else self.dynamic_)[name] = obj

    def del_object(self, name):
        if name in self.objects_:
            del self.objects_[name]
        elif name in self.auxObjects_:
            del self.auxObjects_[name]
        else:
            raise Exception("%s not found" % name)
        if name in self.static_:
            del self.static_[name]
        elif name in self.dynamic_:
            del self.dynamic_[name]

    def del_all_dynamic(self):
        for name in list(self.dynamic_.keys()):
            self.del_object(name)
        self.count_["ball"] = 0

    def get_object_position(self, objName):
        return self.objects_[objName].get_position()

    def set_object_position(self, objName, newPos):
        assert objName in self.dynamic_
        self.dynamic_[objName].set_position(newPos)

    def increment_object_position(self, objName, deltaPos):
        assert objName in self.dynamic_
        self.dynamic_[objName].set_position(self.dynamic_[objName].get_position() + deltaPos)

    def save_to_file(self, outName):
        with open(outName, "wb") as fh:
            pickle.dump({k: copy.deepcopy(v) for k, v in self.objects_.items()}, fh, pickle.HIGHEST_PROTOCOL)

    def draw_arrow(self, pos, direction, fColor=Color(0.0, 0.0, 0.0)):
        raise NotImplementedError("debug arrow overlay (primitives.py:984-988) is outside the simulate-and-render path")

    # -- device ----------------------------------------------------------------------------------
    def _device(self, dt=None, g=None, event_capacity=None):
        """The device mirror, rebuilt when the object layout or dt changed."""
        if dt is None:
            dt, g, event_capacity = self._dev_params
        sig = _DeviceWorld.signature(self, dt, self.dtype_, self.device_)
        if self._dev is None or self._dev.sig != sig:
            if self._dev is not None:
                # the Dynamics that steps this world keeps the event log: fetch what the old mirror still holds
                owner = getattr(self, "_dev_owner", None)
                owner = owner() if owner is not None else None
                if owner is not None:
                    owner._drain_events()
                self._dev.close()
            self._dev = _DeviceWorld(self, dt, g, self.dtype_, self.device_, event_capacity)
            self._dev_params = (dt, g, event_capacity)
        else:
            self._dev.set_g(g)
            self._dev.push(self)
        return self._dev

    def generate_image(self, returnContext=False, cropObject=None, cropSz=None):
        """White canvas + every object in insertion order -> new (ySz, xSz, 4) uint8 array
        (primitives.py:991-1008; the crop arguments are ignored there too)."""
        dev = self._device()
        dev.ensure_styles(self)
        im = dev.batch.render().cpu().numpy()[0]
        return (im, None) if returnContext else im

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_dev"] = None
        d.pop("_dev_owner", None)
        return d


# ------------------------------------------------------------------------------------------
# Dynamics
# ------------------------------------------------------------------------------------------
class Dynamics(object):
    """Event-driven stepper (primitives.py:535-805), executed by the advance kernel."""

    EVENT_CAPACITY = 1 << 16
    EVENT_DRAIN_EVERY = 64      # steps between drains of the device event buffer (one world: << 1000 events per step)

    def __init__(self, world, g=0, deltaT=0.01):
        self.world_ = world
        self.g_ = gm.Point(0, g)
        self.deltaT_ = deltaT
        self._requeue = False
        self._nsteps = 0
        self._events = []      # (step, ball name, partner name) fetched so far
        for name in self.get_dynamic_object_names():
            self._include_object(name)

    # -- bookkeeping -------------------------------------------------------------------------------
    def _include_object(self, name):
        """tCol_ = 0, objCol_ = None, queued (primitives.py:556-563).  On the device this is what
        (re)building the mirror does for every ball, which is how the reference always uses it."""
        assert name in self.get_dynamic_object_names()
        self._force_rebuild = True

    def _dev(self):
        w = self.world_
        if getattr(self, "_force_rebuild", False) and w._dev is not None:
            self._drain_events()
            w._dev.close()
            w._dev = None
        self._force_rebuild = False
        w._dev_owner = weakref.ref(self)
        dev = w._device(self.deltaT_, self.g_, self.EVENT_CAPACITY)
        if self._requeue:
            dev.batch.requeue()
            self._requeue = False
        return dev

    def _drain_events(self):
        dev = self.world_._dev
        if dev is None:
            return
        ev, total = dev.batch.read_events()
        if total > len(ev):
            warnings.warn("collision event log overflowed: %d of %d events since the last drain were kept"
                          % (len(ev), total), RuntimeWarning)
        names = dev.wall_names + dev.ball_names
        base = getattr(dev, "step_base", 0)
        for _, step, ball, partner in ev:
            self._events.append((base + int(step), dev.ball_names[ball], names[partner]))
        dev.batch.clear_events()

    # -- reference API -------------------------------------------------------------------------------
    def set_g(self, g):
        self.g_ = g

    def set_object_velocity(self, objName, vel):
        assert objName in self.get_dynamic_object_names()
        self.get_object(objName).set_velocity(vel)
        self.add_all_dynamic_collision_queue()

    def apply_force(self, objName, force, forceT=None):
        """vel += 0.5 * forceT^2 * F / m (primitives.py:580-594), caches untouched."""
        if forceT is None:
            forceT = self.deltaT_
        obj = self.get_object(objName)
        mass = obj.get_mass()
        assert mass > 0, "Mass has to be a positive number"
        a = gm.Point.from_self(force)
        a.scale(1.0 / mass)
        obj.set_velocity(obj.get_velocity() + 0.5 * forceT * forceT * a)

    def add_all_dynamic_collision_queue(self):
        self._requeue = True

    def get_dynamic_object_names(self):
        return self.world_.get_dynamic_object_names()

    def get_object(self, name):
        return self.world_.get_object(name)

    def get_object_position(self, name):
        return self.world_.get_object_position(name)

    def step(self):
        """One ``deltaT`` of the world (primitives.py:628-656)."""
        dev = self._dev()
        if not hasattr(dev, "step_base"):
            dev.step_base = self._nsteps
        dev.batch.step(1)
        dev.steps += 1
        self._nsteps += 1
        if dev.steps % self.EVENT_DRAIN_EVERY == 0:
            self._drain_events()      # the device keeps EVENT_CAPACITY records; move them to the host log in time
        status, at = dev.batch.get_status()
        dev.pull(self.world_)
        if status[0] != 0:
            raise _status_error(status[0], int(at[0]) + dev.step_base)

    def step_many(self, n, want_frames=True):
        """``n`` x (step() + generate_image()) in ONE device call (extension; the look-ahead window of
        DynamicsHorizon is filled with it): returns the positions after every step, (n, n_balls, 2) float64 in
        dynamic-object order, and the frames (n, ySz, xSz, 4) uint8 (None without want_frames).  The host objects are
        left at the state after the last step; a reference assertion raises like step() does."""
        dev = self._dev()
        if not hasattr(dev, "step_base"):
            dev.step_base = self._nsteps
        if want_frames:
            dev.ensure_styles(self.world_)
        traj, frames = dev.batch.step(int(n), record_traj=True, render_every=1 if want_frames else 0)
        dev.steps += int(n)
        self._nsteps += int(n)
        self._drain_events()
        status, at = dev.batch.get_status()
        dev.pull(self.world_)
        if status[0] != 0:
            raise _status_error(status[0], int(at[0]) + dev.step_base)
        traj = traj[:, 0].cpu().numpy().astype(np.float64)
        return traj, (frames[:, 0].cpu().numpy() if want_frames else None)

    def generate_image(self):
        self._dev()
        return self.world_.generate_image()

    def add_object(self, obj, **kwargs):
        self.world_.add_object(obj, **kwargs)
        for name in self.get_dynamic_object_names():
            self._include_object(name)

    def del_object(self, name):
        self.world_.del_object(name)
        self._force_rebuild = True
        self.add_all_dynamic_collision_queue()

    def del_all_dynamic_objects(self):
        for name in list(self.get_dynamic_object_names()):
            self.del_object(name)
        self.world_.del_all_dynamic()

    def reset_dynamics(self):
        for name in self.get_dynamic_object_names():
            self._include_object(name)
            self.set_object_velocity(name, gm.Point(0, 0))

    # -- collision caches, read back from the device ----------------------------------------------------
    def _cache(self):
        dev = self._dev()
        tcol, partner = dev.batch.get_collision_cache()
        names = dev.wall_names + dev.ball_names
        t = co.OrderedDict((n, float(tcol[0, j])) for j, n in enumerate(dev.ball_names))
        o = co.OrderedDict()
        for j, n in enumerate(dev.ball_names):
            p = int(partner[0, j])
            o[n] = (None, None) if p < 0 else (self.world_.objects_[names[p]], names[p])
        return t, o

    @property
    def tCol_(self):
        return self._cache()[0]

    @property
    def objCol_(self):
        return self._cache()[1]

    def get_colliding_object_names(self):
        """Names with 0 < tCol_ < deltaT (primitives.py:793-805)."""
        return [n for n, t in self.tCol_.items() if 0 < t < self.deltaT_]

    # -- extension: the collision event log -------------------------------------------------------------
    def get_events(self):
        """[(step, ball name, partner name)]: one entry per ``resolve_collision`` call
        (primitives.py:659), in call order; step = number of completed step() calls before it."""
        self._drain_events()
        return list(self._events)


class DynamicsHorizon(object):
    """Rolling look-ahead window (primitives.py:809-852): frames are delivered ``lookAhead`` steps
    late together with the per-step displacements of every ball over that horizon.

    The reference's version cannot run (its ``step()`` returns None and is appended as the image,
    primitives.py:832-836, and ``get_data`` indexes with the ball index instead of the horizon
    index, :848); this one implements the evident intent with the same constructor and
    ``get_data() -> (image, (n_balls, 2*lookAhead) float32)`` signature.

    The window is filled ``chunk`` steps at a time (default: ``lookAhead``): one device call steps and draws the
    chunk (``Dynamics.step_many``) and one launch turns the recorded positions into the look-ahead matrices of all
    its frames (``ppb_horizon_async``), instead of one step + one frame + a Python loop per ``get_data``.  The model
    therefore runs up to ``lookAhead + chunk - 1`` steps ahead of the frame being delivered.
    """

    def __init__(self, model, lookAhead=20, chunk=None):
        self.N_ = lookAhead
        self.model_ = model
        self.names_ = list(model.get_dynamic_object_names())
        self._chunk = max(1, int(chunk if chunk is not None else lookAhead))
        p0 = [model.get_object_position(n) for n in self.names_]
        self._rows = [np.array([[p.x(), p.y()] for p in p0], np.float64)]   # positions after step 0, 1, 2, ...
        self._row0 = 0                  # step index of self._rows[0]
        self._next = 0                  # get_data calls served so far
        self.im_ = co.deque()
        self._mats = co.deque()
        self.outMat_ = np.zeros((len(self.names_), self.N_ * 2), np.float32)
        self._advance(lookAhead)

    def _advance(self, n):
        traj, frames = self.model_.step_many(n)
        self._rows.extend(traj[i] for i in range(n))
        self.im_.extend(frames[i] for i in range(n))

    def step(self):
        """One more step of the window (kept for the reference's method name): returns its frame."""
        self._advance(1)
        return self.im_[-1]

    def _fill_mats(self):
        import torch
        have = self._row0 + len(self._rows) - 1                 # last step recorded
        if have < self._next + self.N_:
            self._advance(max(self._chunk, self._next + self.N_ - have))
        have = self._row0 + len(self._rows) - 1
        first = self._next - self._row0
        window = np.stack(self._rows[first:])                   # rows of steps _next .. have
        dev = self.model_._dev()
        t = torch.from_numpy(window[:, None]).to("cuda:%d" % dev.batch.device)
        if dev.batch.dtype in ("f32", "float32", "fp32"):
            t = t.float()
        mats = dev.batch.horizon(t.contiguous(), self.N_).cpu().numpy()[:, 0]   # (have - _next - N + 1, n_balls, 2N)
        self._mats.extend(mats[i] for i in range(mats.shape[0]))
        del self._rows[:first]
        self._row0 = self._next

    def get_data(self):
        if not self._mats:
            self._fill_mats()
        im = self.im_.popleft()
        self.outMat_[...] = self._mats.popleft()
        self._next += 1
        return im.copy(), self.outMat_.copy()


# ------------------------------------------------------------------------------------------
# sprite makers (primitives.py:33-77) -- the ball sprite is built on the device.  get_block_im /
# get_arrow_im (primitives.py:80-174) are private helpers of the Cairo renderer with no counterpart:
# the render kernel fills wall quads directly.
# ------------------------------------------------------------------------------------------
def get_ball_im(radius=40, fColor=Color(0.0, 0.0, 1.0), sThick=2, sColor=None):
    """-> (None, (2(r+sT), 2(r+sT), 4) uint8): premultiplied ring + disc sprite, channels [r, g, b, a]."""
    from .engine import WorldBatch
    sc = (0.0, 0.0, 0.0, 1.0) if sColor is None else sColor.rgba()
    wb = WorldBatch(1, 1, 0, dtype="f32", canvas=(4, 4))
    try:
        wb.set_styles([(int(sThick), fColor.rgba(), sc)], [], int(radius), int(radius))
        im = wb.read_sprite(0, int(radius))
    finally:
        wb.close()
    return None, im


def _solid_premultiplied(h, w, rgba):
    from .engine import premul_rgba8
    return np.broadcast_to(np.array(premul_rgba8(rgba), np.uint8), (h, w, 4)).copy()


def get_rectangle_im(sz=gm.Point(4, 100), fColor=Color(1.0, 0.0, 0.0)):
    """-> (None, solid (sz.y, sz.x, 4) uint8 sprite) (primitives.py:64-77)."""
    return None, _solid_premultiplied(sz.y_asint(), sz.x_asint(), fColor.rgba())
