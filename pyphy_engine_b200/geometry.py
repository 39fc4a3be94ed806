"""2-D geometry value classes with the reference's ``geometry.py`` interface.

Drop-in for the reference module (``gm.Point``, ``gm.Line``, ``gm.Circle``,
``gm.Bbox``, ``gm.theta2dir``, ``gm.TOL``): same names, same argument meaning,
same return conventions (``None`` for "no intersection", ``np.inf`` for "never").
These are host-side helpers used to *describe* worlds (cages, ball positions,
forces); the per-step arithmetic runs in the CUDA kernels
(``csrc/ppb_math.cuh``), which restate the same formulas.

Every floating-point expression below keeps the reference's operation order, so
a cage built here has bit-identical vertices to one built by the reference
(the golden fixtures under ``tests/golden`` depend on that).  Reference line
numbers are cited per method (``geometry.py:<line>``).
"""
from __future__ import annotations

import math

import numpy as np

TOL = 1e-08  # geometry.py:7

__all__ = ["TOL", "Point", "Line", "Circle", "Bbox", "theta2dir"]


def _round_half_away(v: float) -> int:
    """Python-2 ``round``: halves go away from zero (the reference ran on py2; geometry.py:25-29)."""
    return int(math.floor(abs(v) + 0.5)) * (-1 if v < 0 else 1)


class Point(object):
    """A 2-vector of Python floats (geometry.py:9-175).

    ``Point +/- scalar`` moves the point along its own direction by that amount
    (geometry.py:89-93, 101-105); ``Point * scalar`` scales it.
    """

    __slots__ = ("x_", "y_")

    def __init__(self, x=0, y=0):
        self.x_ = float(x)
        self.y_ = float(y)

    @classmethod
    def from_self(cls, other):
        return cls(x=other.x(), y=other.y())

    # -- accessors ---------------------------------------------------------------------------
    def x(self):
        return self.x_

    def y(self):
        return self.y_

    def x_asint(self):
        return _round_half_away(self.x_)

    def y_asint(self):
        return _round_half_away(self.y_)

    # -- in-place operations -----------------------------------------------------------------
    def scale(self, xScale, yScale=None):
        """geometry.py:34-39 (multiplies in place; a single factor applies to both axes)."""
        self.x_ = xScale * self.x_
        self.y_ = (xScale if yScale is None else yScale) * self.y_

    def make_unit_norm(self):
        """geometry.py:48-51: multiply by 1.0/mag; a zero vector stays zero."""
        m = self.mag()
        if m > 0:
            self.scale(1.0 / m)

    # -- pure functions ----------------------------------------------------------------------
    def mag(self):
        return math.sqrt(self.x_ * self.x_ + self.y_ * self.y_)

    def get_scaled_vector(self, scale):
        out = Point(self.x_, self.y_)
        out.make_unit_norm()
        out.scale(scale)
        return out

    def dot(self, other):
        return self.x_ * other.x() + self.y_ * other.y()

    def project(self, other):
        """Component of self along ``other`` (geometry.py:58-63)."""
        u = Point(other.x(), other.y())
        u.make_unit_norm()
        u.scale(self.dot(u))
        return u

    def cosine(self, other):
        a, b = Point(self.x_, self.y_), Point(other.x(), other.y())
        a.make_unit_norm()
        b.make_unit_norm()
        return a.dot(b)

    def reflect_normal(self, other):
        """Reflect ``other`` about the normal ``self`` (geometry.py:74-82): -parallel + orthogonal."""
        n = Point(self.x_, self.y_)
        n.make_unit_norm()
        par = other.project(n)
        orth = other - par
        par.scale(-1)
        return par + orth

    def _along_self(self, amount):
        u = Point(self.x_, self.y_)
        u.make_unit_norm()
        u.scale(amount)
        return u

    def __add__(self, other):
        if isinstance(other, Point):
            return Point(self.x_ + other.x_, self.y_ + other.y_)
        return self + self._along_self(other)

    def __sub__(self, other):
        if isinstance(other, Point):
            return Point(self.x_ - other.x_, self.y_ - other.y_)
        return self - self._along_self(other)

    def __mul__(self, scale):
        return Point(self.x_ * scale, self.y_ * scale)

    __rmul__ = __mul__

    def __str__(self):
        return "(%.2f, %.2f)" % (self.x_, self.y_)

    def __repr__(self):
        return "Point(%r, %r)" % (self.x_, self.y_)

    # quadrant tests with self as the origin (geometry.py:120-133)
    def is_quad1(self, pt):
        return pt.x() >= self.x_ and pt.y() >= self.y_

    def is_quad2(self, pt):
        return pt.x() <= self.x_ and pt.y() >= self.y_

    def is_quad3(self, pt):
        return pt.x() <= self.x_ and pt.y() <= self.y_

    def is_quad4(self, pt):
        return pt.x() >= self.x_ and pt.y() <= self.y_

    def distance(self, pt, distType="L2"):
        if distType != "L2":
            raise Exception("DistType: %s not recognized" % distType)
        return (self - pt).mag()

    def get_angle(self, isRadian=False):
        rad = math.atan2(self.y_, self.x_)
        return rad if isRadian else math.degrees(rad)

    def rotate_point(self, rot, isRadian=False):
        """geometry.py:153-164 (degrees unless isRadian)."""
        m = self.mag()
        theta = self.get_angle(isRadian=isRadian) + rot
        rad = theta if isRadian else math.radians(theta)
        out = Point(np.cos(rad), np.sin(rad))
        out.make_unit_norm()
        out.scale(m)
        return out

    def get_angle_between(self, other):
        """geometry.py:166-175: acos of the clamped cosine (raises ValueError above 1 + 1e-6)."""
        c = self.cosine(other)
        if c < 1 + 1e-6:
            c = min(1, c)
        return math.acos(c)


def theta2dir(theta):
    """Unit vector at ``theta`` degrees from the x axis (geometry.py:177-188)."""
    assert -180 < theta <= 180
    rad = np.pi * (theta / 180.0)
    pt = Point(np.cos(rad), np.sin(rad))
    pt.make_unit_norm()
    return pt


class Line(object):
    """Directed line st -> en with canonical form a x + b y + c = 0 (geometry.py:191-357).

    The coefficients are divided by ``|a|`` (not the norm of (a, b)) unless ``|a| <= TOL``,
    in which case ``a`` is set to 0 and b, c stay unscaled (geometry.py:203-216).
    """

    def __init__(self, pt1, pt2):
        self.st_ = pt1
        self.en_ = pt2
        self.make_canonical()

    @classmethod
    def from_self(cls, other):
        return cls(other.st(), other.en())

    def make_canonical(self):
        s, e = self.st_, self.en_
        a = float(-(e.y() - s.y()))
        b = float(e.x() - s.x())
        c = float(s.x() * e.y() - s.y() * e.x())
        am = abs(a)
        if am > TOL:
            a, b, c = a / am, b / am, c / am
        else:
            a = 0.0
        self.a_, self.b_, self.c_ = a, b, c

    def a(self):
        return self.a_

    def b(self):
        return self.b_

    def c(self):
        return self.c_

    def st(self):
        return Point(self.st_.x(), self.st_.y())

    def mutable_st(self):
        return self.st_

    def en(self):
        return Point(self.en_.x(), self.en_.y())

    def mutable_en(self):
        return self.en_

    def get_direction(self):
        d = self.en_ - self.st_
        d.make_unit_norm()
        return d

    def distance_to_point(self, pt):
        """Signed distance (geometry.py:244-248)."""
        num = self.a_ * pt.x() + self.b_ * pt.y() + self.c_
        den = np.sqrt(np.power(self.a_, 2) + np.power(self.b_, 2))
        return num / den

    def get_normal(self):
        n = Point(-self.a_, -self.b_)
        n.make_unit_norm()
        return n

    def get_normal_towards_point(self, pt):
        """The unit normal on whose side ``pt`` lies (geometry.py:258-266)."""
        n = self.get_normal()
        if self.get_intersection_ray(Line(pt, pt + n)) is None:
            return n
        n.scale(-1)
        return n

    def __str__(self):
        return "(%.2f, %.2f, %.2f)" % (self.a_, self.b_, self.c_)

    def get_point_location(self, pt, tol=TOL):
        """+1 / -1 / 0: sign of a x + b y + c with a dead band of ``tol``."""
        v = self.a_ * pt.x() + self.b_ * pt.y() + self.c_
        return 1 if v > tol else (-1 if v < -tol else 0)

    def get_relative_location_points(self, pt1, pt2, tol=TOL):
        """+1 if pt2 lies from pt1 along the line direction, -1 if against it, else 0."""
        d = pt2 - pt1
        d.make_unit_norm()
        cos = d.cosine(self.get_direction())
        if 1 - tol < cos < 1 + tol:
            return 1
        if -1 - tol < cos < -1 + tol:
            return -1
        return 0

    def is_on_segment(self, pt, tol=TOL):
        total = self.st_.distance(pt) + self.en_.distance(pt)
        span = self.st_.distance(self.en_)
        return bool(span - tol <= total <= span + tol)

    def get_point_along_line(self, pt, distance):
        d = self.get_direction()
        d.scale(distance)
        return pt + d

    def get_intersection(self, l2):
        """Point common to the two infinite lines, None when parallel (geometry.py:327-345)."""
        nr = l2.a() * self.c_ - self.a_ * l2.c()
        dr = self.a_ * l2.b() - l2.a() * self.b_
        if dr == 0:
            return None
        y = nr / dr
        if self.a_ == 0:
            x = -(l2.c() + l2.b() * y) / l2.a()
        else:
            x = -(self.c_ + self.b_ * y) / self.a_
        return Point(x, y)

    def get_intersection_ray(self, l2):
        """As get_intersection, with ``l2`` a ray from l2.st() through l2.en()."""
        pt = self.get_intersection(l2)
        if pt is not None and l2.get_relative_location_points(l2.st(), pt) != 1:
            return None
        return pt


class Circle(object):
    """geometry.py:361-466."""

    def __init__(self, radius=20, center=None):
        self.r_ = radius
        self.c_ = Point(0, 0) if center is None else center

    def get_contact_point_pseudo_tangent(self, l):
        """Contact point of the tangent parallel to ``l`` on l's side (geometry.py:369-390)."""
        toward = l.get_normal()
        toward.scale(self.r_)
        for _ in range(2):
            pt = self.c_ + toward
            if l.get_intersection_ray(Line(self.c_, pt)) is not None:
                return pt
            toward.scale(-1.0)
        raise AssertionError("the line does not face the circle from either side")

    def is_intersect_line(self, l):
        return bool(np.abs(l.distance_to_point(self.c_)) <= self.r_)

    def intersect_moving_circle(self, circ2, v21):
        """Time until ``circ2`` (moving at ``v21`` relative to self) touches self.

        Returns ``(tCol, centre of circ2 at contact, tangent direction)`` or
        ``(inf, None, None)`` (geometry.py:400-466).  Law-of-sines construction, including the
        [R-0.1, R] overlap nudge and the reference's assertions.
        """
        p1 = self.c_
        origin = p1 - p1
        rel = circ2.c_ - p1
        R = self.r_ + circ2.r_
        d = rel.distance(origin)
        if (R - 0.1) <= d <= R:
            d = (R - d) + d + 1e-6
        assert d >= R, "dball: %f, R: %f" % (d, R)
        if not Circle(R, origin).is_intersect_line(Line(rel, rel + v21)):
            return np.inf, None, None
        back = rel - origin
        back.scale(-1)
        theta = back.get_angle_between(v21)
        assert np.abs(theta) < np.pi / 2
        theta = np.abs(theta)
        if theta == 0:
            dist = d - R
        else:
            sin_c = d / (R / np.sin(theta))
            theta_c = np.pi - math.asin(sin_c)
            theta_a = np.pi - (theta_c + theta)
            assert theta_a >= 0
            dist = (R / np.sin(theta)) * np.sin(theta_a)
        speed = v21.mag()
        if speed == 0:
            return np.inf, None, None
        vdir = Point.from_self(v21)
        vdir.make_unit_norm()
        t_col = dist / speed
        c2 = rel + dist * vdir
        tangent = Line(c2, origin).get_normal()
        return t_col, c2 + p1, tangent


class Bbox(object):
    """A closed 4-vertex polygon and its four edge lines (geometry.py:471-613)."""

    def __init__(self, lTop, lBot, rBot, rTop):
        self.vert_ = [lTop, lBot, rBot, rTop]
        self.N_ = 4
        self.lines_ = []
        self.make_lines()

    @classmethod
    def from_list(cls, pts):
        assert len(pts) == 4
        return cls(*[Point(p.x(), p.y()) for p in pts])

    def move(self, offset):
        self.vert_ = [v + offset for v in self.vert_]
        self.make_lines()

    def make_lines(self):
        n = self.N_
        self.lines_ = [Line(self.vert_[i], self.vert_[(i + 1) % n]) for i in range(n)]

    def get_lines(self):
        return self.lines_

    def is_point_inside(self, pt):
        sides = [l.get_point_location(pt) for l in self.lines_]
        return all(s >= 0 for s in sides) or all(s <= 0 for s in sides)

    def is_intersect_line(self, los):
        """True unless all four vertices lie strictly on one side of ``los``."""
        sides = [los.get_point_location(v) for v in self.vert_]
        return not all(s == sides[0] for s in sides[1:])

    def is_intersect_line_ray(self, los):
        return self.get_intersection_with_line_ray(los)[0] is not None

    def find_closest_interior_point(self, srcPt, pts, getIndex=False):
        best, best_d, best_i = None, np.inf, None
        for i, pt in enumerate(pts):
            if pt is None or not self.is_point_inside(pt):
                continue
            d = srcPt.distance(pt)
            if d < best_d:
                best, best_d, best_i = pt, d, i
        return (best, best_d, best_i) if getIndex else (best, best_d)

    def _edges(self):
        n = self.N_
        return [Line(self.vert_[i], self.vert_[(i + 1) % n]) for i in range(n)]

    def get_intersection_with_line(self, l):
        """Closest (to l.st()) crossing of the infinite line ``l`` with the polygon: (point, distance)."""
        pts = [l.get_intersection(e) for e in self._edges()]
        pt, dist, _ = self.find_closest_interior_point(l.st(), pts, getIndex=True)
        return pt, dist

    def get_line_of_first_intersection_with_ray(self, l):
        return None

    def get_intersection_with_line_ray(self, l):
        pts = [e.get_intersection_ray(l) for e in self._edges()]
        return self.find_closest_interior_point(l.st(), pts)

    def get_toc_with_bbox(self, bbox, vel):
        raise Exception("This function is not ready")
