// ppb_math.cuh -- time-of-collision and response arithmetic.
//
// Two formulations of the same algorithm:
//   Compat<double> : every expression keeps the reference's operation order
//                    (geometry.py / dynamics.py), compiled with -fmad=false, trig from
//                    portable_trig.h -> bit-identical with an IEEE fp64 CPU evaluation.
//   Fast<float>    : the same decisions (normal towards the line, speed <= 0 rejection,
//                    on-segment test, overlap nudge, smaller root of the contact
//                    quadratic, centre-line impulse) re-derived for fp32 throughput.
#pragma once
#include "portable_trig.h"
#include "ppb_common.cuh"

namespace ppb {

#define PPB_TOL 1e-08 /* geometry.py:7 */

// result of one (ball, object) test
template <typename T>
struct Toc {
  T t;        // time of collision, +inf when none
  V2<T> n;    // wall: stored normal (-n towards the line, dynamics.py:57-58); ball: contact tangent
  V2<T> fv;   // ball-ball only: post-collision velocity of ball 1 (dynamics.py:125-129)
  int has_fv; // the ball-ball test got as far as writing futureVel_
  int err;    // PPB_ST_* when the reference would have raised
};

// =============================== Compat (fp64) ===============================
struct Compat {
  typedef double T;
  typedef V2<double> P;

  __device__ __forceinline__ static T mag(P p) { return sqrt(p.x * p.x + p.y * p.y); }  // geometry.py:31
  __device__ __forceinline__ static P scale(P p, T s) { return mk<T>(s * p.x, s * p.y); }  // :34
  __device__ __forceinline__ static P unit(P p) {  // :48 (multiplies by 1.0/mag; zero stays zero)
    T m = mag(p);
    if (m > 0) return scale(p, 1.0 / m);
    return p;
  }
  __device__ __forceinline__ static P project(P self, P other) {  // :58
    P u = unit(other);
    T pm = dot(self, u);
    return scale(u, pm);
  }
  __device__ __forceinline__ static T cosine(P a, P b) { return dot(unit(a), unit(b)); }  // :66
  __device__ __forceinline__ static P reflect(P nrml, P v) {  // Point.reflect_normal :74
    P s = unit(nrml);
    P prll = project(v, s);
    P orth = v - prll;
    prll = scale(prll, -1.0);
    return prll + orth;
  }

  struct Line {
    P st, en;
    T a, b, c;
  };
  __device__ __forceinline__ static Line line(P p1, P p2) {  // Line.make_canonical :203
    Line l;
    l.st = p1;
    l.en = p2;
    l.a = -(p2.y - p1.y);
    l.b = p2.x - p1.x;
    l.c = p1.x * p2.y - p1.y * p2.x;
    T am = fabs(l.a);
    if (am > PPB_TOL) {
      l.a = l.a / am;
      l.b = l.b / am;
      l.c = l.c / am;
    } else {
      l.a = 0.0;
    }
    return l;
  }
  __device__ __forceinline__ static P normal(const Line &l) { return unit(mk<T>(-l.a, -l.b)); }  // :251
  __device__ __forceinline__ static bool on_segment(const Line &l, P p) {  // :310
    T d1 = mag(l.st - p), d2 = mag(l.en - p), d3 = mag(l.st - l.en);
    T ds = d1 + d2;
    return (ds <= d3 + PPB_TOL) && (ds >= d3 - PPB_TOL);
  }
  // Line.get_intersection_ray(ray) :327-357 with ray = Line(o, o + dir)
  __device__ __forceinline__ static bool hit_ray(const Line &s, P o, P dir, P &out) {
    Line r = line(o, o + dir);
    T nr = r.a * s.c - s.a * r.c;
    T dr = s.a * r.b - r.a * s.b;
    if (dr == 0) return false;
    T y = nr / dr, x;
    if (s.a == 0) x = -(r.c + r.b * y) / r.a;
    else x = -(s.c + s.b * y) / s.a;
    P p = mk<T>(x, y);
    // get_relative_location_points(ray.st, p) == 1   (:288-307)
    P pd = unit(p - r.st);
    P ld = unit(r.en - r.st);
    T c = cosine(pd, ld);
    if (!(c > 1 - PPB_TOL && c < 1 + PPB_TOL)) return false;
    out = p;
    return true;
  }

  // one edge of dynamics.get_toc_ball_wall (dynamics.py:20-59)
  __device__ static Toc<T> edge(P pos, P vel, T r, P v0, P v1) {
    Toc<T> o;
    o.t = Lim<T>::inf();
    o.err = 0;
    o.has_fv = 0;
    o.n = mk<T>(0, 0);
    o.fv = o.n;
    Line l = line(v0, v1);
    // get_normal_towards_point (:258-266) followed by scale(-1) (dynamics.py:22-23)
    P n0 = normal(l);
    P ip;
    P nrml = hit_ray(l, pos, n0, ip) ? scale(n0, -1.0) : n0;
    nrml = scale(nrml, -1.0);
    T speed = dot(vel, nrml);
    if (speed <= 0) return o;
    P velOrth = vel - nrml * speed;
    if (!hit_ray(l, pos, nrml, ip)) {  // dynamics.py:34
      o.err = PPB_ST_ASSERT_INTPOINT;
      return o;
    }
    T distCenter = mag(pos - ip);
    T dist = distCenter - r;
    if (dist < 0) {  // dynamics.py:38-47
      if (on_segment(l, ip)) o.err = PPB_ST_TRAP_AMISS;
      return o;
    }
    T t = dist / speed;
    P lp = ip + velOrth * t;
    if (!on_segment(l, lp)) return o;
    o.t = t;
    o.n = scale(nrml, -1.0);
    return o;
  }

  // dynamics.get_toc_ball_wall: the four edges in Bbox order, strict '<' (dynamics.py:17-63)
  __device__ static void wall(P pos, P vel, T r, const T *wv, T &tw, P &nw, int &err) {
    tw = Lim<T>::inf();
    nw = mk<T>(0, 0);
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
      const int e1 = (e + 1) & 3;
      const Toc<T> o = edge(pos, vel, r, mk<T>(wv[2 * e], wv[2 * e + 1]), mk<T>(wv[2 * e1], wv[2 * e1 + 1]));
      if (o.err) {
        err = o.err;
        return;
      }
      if (o.t < tw) {
        tw = o.t;
        nw = o.n;
      }
    }
  }

  // one scan item of the collide kernel = one wall edge, staged in shared memory
  struct Edge {
    P v0, v1;
  };
  __device__ __forceinline__ static Edge make_edge(P v0, P v1) {
    Edge e;
    e.v0 = v0;
    e.v1 = v1;
    return e;
  }
  // returns the PPB_ST_* the reference would raise (0 = none); t = +inf when the edge is not hit
  __device__ __forceinline__ static int edge_toc(P pos, P vel, T r, const Edge &e, T &t, P &n) {
    const Toc<T> o = edge(pos, vel, r, e.v0, e.v1);
    t = o.t;
    n = o.n;
    return o.err;
  }

  // dynamics.get_toc_ball_ball (dynamics.py:67-136) + Circle.intersect_moving_circle
  // (geometry.py:400-466)
  __device__ static Toc<T> ball(P pos1, P vel1, T r1, T m1, P pos2, P vel2, T r2, T m2) {
    Toc<T> o;
    o.t = Lim<T>::inf();
    o.err = 0;
    o.has_fv = 0;
    o.n = mk<T>(0, 0);
    o.fv = o.n;
    P relVel = vel2 - vel1;
    P colDir = unit(pos2 - pos1);
    T speed = -dot(relVel, colDir);
    if (speed <= 0) return o;
    P p11 = pos1 - pos1;
    P p21 = pos2 - pos1;
    T R = r1 + r2;
    T dBall = mag(p21 - p11);
    if (dBall >= (R - 0.1) && dBall <= R) dBall = (R - dBall) + dBall + 1e-6;
    if (!(dBall >= R)) {
      o.err = PPB_ST_ASSERT_OVERLAP;
      return o;
    }
    Line lv = line(p21, p21 + relVel);
    // Circle.is_intersect_line -> Line.distance_to_point (geometry.py:244-248, 392-398)
    T dl = lv.a * p11.x + lv.b * p11.y + lv.c;
    dl = fabs(dl / sqrt(lv.a * lv.a + lv.b * lv.b));
    if (!(dl <= R)) return o;
    P c1c2 = scale(p21 - p11, -1.0);
    // get_angle_between (geometry.py:166-175)
    T ct = dot(unit(c1c2), unit(relVel));
    if (ct < 1 + 1e-6) ct = (1 < ct) ? 1.0 : ct;
    if (!(ct >= -1.0 && ct <= 1.0)) {
      o.err = PPB_ST_MATH_DOMAIN;
      return o;
    }
    T theta = ptrig_acos(ct);
    const T kPi = 3.141592653589793;  // np.pi
    if (!(fabs(theta) < kPi / 2)) {
      o.err = PPB_ST_ASSERT_THETA;
      return o;
    }
    theta = fabs(theta);
    T dist;
    if (theta == 0) {
      dist = dBall - R;
    } else {
      T sinC = dBall / (R / ptrig_sin(theta));
      if (!(sinC >= -1.0 && sinC <= 1.0)) {
        o.err = PPB_ST_MATH_DOMAIN;
        return o;
      }
      T thetaC = kPi - ptrig_asin(sinC);
      T thetaA = kPi - (thetaC + theta);
      if (!(thetaA >= 0)) {
        o.err = PPB_ST_ASSERT_THETA_A;
        return o;
      }
      dist = (R / ptrig_sin(theta)) * ptrig_sin(thetaA);
    }
    T vmag = mag(relVel);
    if (vmag == 0) return o;
    P vDir = unit(relVel);
    T tCol = dist / vmag;
    P newP2 = p21 + vDir * dist;
    Line lc = line(newP2, p11);
    P n = normal(lc);
    // response, dynamics.py:97-129
    P vOth1 = project(vel1, n), vOth2 = project(vel2, n);
    P vCol1 = vel1 - vOth1, vCol2 = vel2 - vOth2;
    P v1n = (vCol1 * (m1 - m2) + vCol2 * (2 * m2)) * (1.0 / (m1 + m2));
    o.fv = v1n + vOth1;
    o.has_fv = 1;
    o.n = n;
    o.t = tCol;
    return o;
  }

  __device__ __forceinline__ static T round_half_away(T x) { return round(x); }
};

// ================================ Fast (fp32) ================================
struct Fast {
  typedef float T;
  typedef V2<float> P;

  __device__ __forceinline__ static P reflect(P nrml, P v) {
    // v - 2 (v.s) s with s = unit(nrml)   (geometry.py:74-82)
    float m2 = nrml.x * nrml.x + nrml.y * nrml.y;
    if (!(m2 > 0.f)) return v;
    float k = 2.f * (v.x * nrml.x + v.y * nrml.y) / m2;
    return mk<T>(v.x - k * nrml.x, v.y - k * nrml.y);
  }

  // dynamics.get_toc_ball_wall (dynamics.py:8-63) for the four edges of one wall, straight-line
  // code so that the four edges overlap in the pipeline.  Quantities are kept multiplied by the
  // edge length L (no normalisation of the edge direction is needed):
  //   s_un = L * signed distance of the centre from the line,  spu = L * (v . n) with n the
  //   unit normal from the ball towards the line,  num = L * (|dist| - r),  t = num / spu,
  //   ufL / ucL = L * (foot / contact point measured along the edge from its first vertex).
  // The returned normal is not normalised (reflect() divides by its squared length).
  __device__ __forceinline__ static void wall(P pos, P vel, T r, const T *wv, T &tw, P &nw, int &err) {
    tw = Lim<T>::inf();
    nw = mk<T>(0.f, 0.f);
    int e_err = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int e1 = (e + 1) & 3;
      const float ex = wv[2 * e1] - wv[2 * e], ey = wv[2 * e1 + 1] - wv[2 * e + 1];
      const float px = pos.x - wv[2 * e], py = pos.y - wv[2 * e + 1];
      const float s_un = px * ey - py * ex;
      const float vn_un = vel.x * ey - vel.y * ex;
      // a centre exactly on the line keeps -get_normal() as the reference does (geometry.py:258-266)
      const float sg = (s_un < 0.f) ? 1.f : -1.f;
      const float spu = sg * vn_un;
      const float L2 = ex * ex + ey * ey;
      const float L = sqrtf(L2);
      const float num = fabsf(s_un) - r * L;
      const float t = __fdividef(num, spu);
      const float ufL = px * ex + py * ey;
      const float ucL = ufL + t * (vel.x * ex + vel.y * ey);
      const float epsL = 1e-4f * L;
      const bool appr = (spu > 0.f) && (L2 > 0.f);              // speed <= 0: dynamics.py:25
      const bool foot_on = (ufL >= -epsL) && (ufL <= L2 + epsL);
      const bool hit_on = (ucL >= -epsL) && (ucL <= L2 + epsL);  // dynamics.py:52
      if (e_err == 0 && appr) {
        if (s_un == 0.f) e_err = PPB_ST_ASSERT_INTPOINT;          // dynamics.py:34
        else if (num < 0.f && foot_on) e_err = PPB_ST_TRAP_AMISS;  // dynamics.py:38-45
      }
      if (e_err == 0 && appr && num >= 0.f && hit_on && t < tw) {
        tw = t;
        nw = mk<T>(-sg * ey, sg * ex);
      }
    }
    if (e_err) err = e_err;
  }

  // one scan item of the collide kernel = one wall edge, staged in shared memory with the
  // quantities that do not depend on the ball
  struct __align__(16) Edge {
    float x0, y0;    // first vertex
    float ex, ey;    // edge vector
    float L, L2;     // its length and squared length
    float epsL;      // on-segment tolerance scaled by L (1e-4 L)
    float pad;
  };
  __device__ __forceinline__ static Edge make_edge(P v0, P v1) {
    Edge e;
    e.x0 = v0.x; e.y0 = v0.y;
    e.ex = v1.x - v0.x; e.ey = v1.y - v0.y;
    e.L2 = e.ex * e.ex + e.ey * e.ey;
    e.L = sqrtf(e.L2);
    e.epsL = 1e-4f * e.L;
    e.pad = 0.f;
    return e;
  }
  // same arithmetic as one iteration of wall() above
  __device__ __forceinline__ static int edge_toc(P pos, P vel, T r, const Edge &e, T &t_out, P &n) {
    const float px = pos.x - e.x0, py = pos.y - e.y0;
    const float s_un = px * e.ey - py * e.ex;
    const float vn_un = vel.x * e.ey - vel.y * e.ex;
    const float sg = (s_un < 0.f) ? 1.f : -1.f;
    const float spu = sg * vn_un;
    const float num = fabsf(s_un) - r * e.L;
    const float t = __fdividef(num, spu);
    const float ufL = px * e.ex + py * e.ey;
    const float ucL = ufL + t * (vel.x * e.ex + vel.y * e.ey);
    const bool appr = (spu > 0.f) && (e.L2 > 0.f);                    // speed <= 0: dynamics.py:25
    const bool foot_on = (ufL >= -e.epsL) && (ufL <= e.L2 + e.epsL);
    const bool hit_on = (ucL >= -e.epsL) && (ucL <= e.L2 + e.epsL);    // dynamics.py:52
    t_out = Lim<T>::inf();
    n = mk<T>(-sg * e.ey, sg * e.ex);
    if (appr) {
      if (s_un == 0.f) return PPB_ST_ASSERT_INTPOINT;          // dynamics.py:34
      if (num < 0.f && foot_on) return PPB_ST_TRAP_AMISS;      // dynamics.py:38-45
      if (num >= 0.f && hit_on) t_out = t;
    }
    return 0;
  }

  // dynamics.py:67-136 + geometry.py:400-466 (smaller root of |p + v t| = R)
  __device__ __forceinline__ static Toc<T> ball(P pos1, P vel1, T r1, T m1, P pos2, P vel2, T r2, T m2) {
    Toc<T> o;
    o.t = Lim<T>::inf();
    o.err = 0;
    o.has_fv = 0;
    o.n = mk<T>(0.f, 0.f);
    o.fv = o.n;
    const float px = pos2.x - pos1.x, py = pos2.y - pos1.y;
    const float vx = vel2.x - vel1.x, vy = vel2.y - vel1.y;
    const float pv = px * vx + py * vy;
    if (!(pv < 0.f)) return o;              // speed = -relVel . colDir <= 0   (dynamics.py:86)
    const float R = r1 + r2;
    const float D2 = px * px + py * py;
    if (D2 < (R - 0.1f) * (R - 0.1f)) {     // geometry.py:413
      o.err = PPB_ST_ASSERT_OVERLAP;
      return o;
    }
    const float vv = vx * vx + vy * vy;
    const float cr = px * vy - py * vx;     // |p x v| = (distance of the path from centre 1) * |v|
    const float disc = R * R * vv - cr * cr;  // = pv^2 - vv (D^2 - R^2)
    if (disc < 0.f) return o;               // geometry.py:416-418 (misses the big circle)
    const float D = sqrtf(D2);
    float gap = D - R;
    if (gap <= 0.f) gap = 1e-6f;            // geometry.py:411-412 nudge
    const float c = gap * (D + R);          // |p|^2 - R^2 without cancellation
    const float t = c / (-pv + sqrtf(disc));
    // contact: centre of ball 2 relative to ball 1 at time t; unit centre line
    float cx = px + vx * t, cy = py + vy * t;
    const float inv = rsqrtf(cx * cx + cy * cy);
    cx *= inv;
    cy *= inv;
    // exchange of the centre-line components (dynamics.py:97-129)
    const float a1 = vel1.x * cx + vel1.y * cy;
    const float a2 = vel2.x * cx + vel2.y * cy;
    const float k = (2.f * m2 / (m1 + m2)) * (a2 - a1);
    o.fv = mk<T>(vel1.x + k * cx, vel1.y + k * cy);
    o.has_fv = 1;
    o.n = mk<T>(-cy, cx);
    o.t = t;
    return o;
  }

  __device__ __forceinline__ static T round_half_away(T x) { return roundf(x); }
};

template <typename T>
struct MathOf;
template <>
struct MathOf<float> {
  typedef Fast type;
};
template <>
struct MathOf<double> {
  typedef Compat type;
};

}  // namespace ppb
