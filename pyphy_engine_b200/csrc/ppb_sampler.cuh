// ppb_sampler.cuh -- device-side sampler of random rectangular tables (sm_100a).
//
// One thread draws one world from the distribution of the reference's DataSaver with isRect=True:
// cage (dataio.py:220-230, 277-297 + primitives.py:177-209), balls by rejection sampling
// (dataio.py:309-379), initial kick (dataio.py:382-416, primitives.py:580-594) -- same loop structure
// and arithmetic (fp64), but on a counter-based Philox4x32-10 stream keyed by (seed, GLOBAL world
// index), so world w is the same whatever the batch split.  It is distribution-equivalent to the
// reference, not stream-equivalent: the stream-exact generator is dataio.DataSaver on the host.
#pragma once
#include "ppb_common.cuh"

namespace ppb {

struct SamplerParams {
  double arena, w_thick, mn_wlen, mx_wlen, mn_ball, mx_ball, mn_force, mx_force, force_t;
  unsigned long long seed;
  long long world_index0;   // global index of the batch's world 0 (sharded jobs)
  int32_t max_tries;
  int32_t n_theta;          // 0: rectangular table; > 0: rhombus with the side angle drawn from theta[] (degrees)
  double theta[4];
};

struct Philox {
  uint32_t key[2], ctr[4], out[4];
  int have;
  __device__ Philox(unsigned long long seed, unsigned long long stream) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = 0u; ctr[1] = 0u; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
    have = 0;
  }
  __device__ void round10() {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    if (++ctr[0] == 0u) ++ctr[1];
  }
  // uniform double in [0, 1) with 53 random bits (the construction numpy's random_sample uses)
  __device__ double rand() {
    if (have == 0) { round10(); have = 2; }
    const uint32_t a = out[4 - 2 * have] >> 5, b = out[5 - 2 * have] >> 6;
    --have;
    return ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
  }
};

struct SLine { double a, b, c; };
// Line.make_canonical (geometry.py:203-216)
__device__ inline SLine s_line(double x0, double y0, double x1, double y1) {
  SLine l;
  l.a = -(y1 - y0); l.b = x1 - x0; l.c = x0 * y1 - y0 * x1;
  const double am = fabs(l.a);
  if (am > 1e-8) { l.a /= am; l.b /= am; l.c /= am; }
  else l.a = 0.0;
  return l;
}
__device__ inline void s_unit(double &x, double &y) {   // Point.make_unit_norm (geometry.py:48-51)
  const double m = sqrt(x * x + y * y);
  if (m > 0) { const double inv = 1.0 / m; x = inv * x; y = inv * y; }
}

template <typename T>
__global__ void sample_rect_worlds_kernel(StatePtrs S, int n_balls, int n_walls, int world0, int count, SamplerParams P,
                                          int32_t step_value, int32_t *fail_count) {
  typedef typename Vec2<T>::type T2;
  typedef typename Vec4<T>::type T4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int w = world0 + i;
  Philox rng(P.seed, (unsigned long long)(P.world_index0 + w));
  const double Tk = P.w_thick;
  double px[4], py[4];
  bool failed = false;
  if (P.n_theta <= 0) {
    // ---- cage: add_rectangular_walls -> _create_walls(xleft, ytop, (0, 90, 180), (hlen - T, vlen, hlen - T)) ----
    const double hlen = floor(P.mn_wlen + rng.rand() * (P.mx_wlen - P.mn_wlen));
    const double vlen = floor(P.mn_wlen + rng.rand() * (P.mx_wlen - P.mn_wlen));
    const double xleft = floor(rng.rand() * (P.arena - (hlen + Tk)));
    const double ytop = floor(rng.rand() * (P.arena - (vlen + Tk)));
    // theta2dir(0 / 90 / 180) as numpy evaluates them (geometry.py:177-188)
    const double dirx[3] = {1.0, 6.123233995736766e-17, -1.0};
    const double diry[3] = {0.0, 1.0, 1.2246467991473532e-16};
    const double wl[3] = {hlen - Tk, vlen, hlen - Tk};
    px[0] = xleft; py[0] = ytop;
    for (int k = 0; k < 3; ++k) { px[k + 1] = px[k] + wl[k] * dirx[k]; py[k + 1] = py[k] + wl[k] * diry[k]; }
  } else {
    // ---- rhombus: sample_walls (dataio.py:243-274) -> _create_walls(xleft, yleft, (-th, th, 180 - th), (h, h, h)) ----
    const double kPiS = 3.141592653589793;
    double wtheta = 0, hlen = 0, xleft = 0, yleft = 0;
    int tries = 0;
    for (;;) {
      // rand_.permutation(n)[0] is uniform over the angles
      int pick = (int)(rng.rand() * P.n_theta);
      if (pick >= P.n_theta) pick = P.n_theta - 1;
      wtheta = P.theta[pick];
      const double rad = (wtheta * kPiS) / 180.0;
      hlen = P.mn_wlen + rng.rand() * (P.mx_wlen - P.mn_wlen);
      const double xlen = (wtheta == 90.0) ? hlen : hlen * cos(rad), ylen = (wtheta == 90.0) ? hlen : hlen * sin(rad);
      const double xleftmin = Tk, xleftmax = P.arena - (2 * xlen + 2 * Tk);
      const double yleftmin = ylen + Tk, yleftmax = (wtheta == 90.0) ? P.arena - Tk : P.arena - (ylen + Tk);
      if (!(xleftmin <= 0 || yleftmin <= 0 || xleftmax < xleftmin || yleftmax < yleftmin)) {
        xleft = xleftmin + floor(rng.rand() * (xleftmax - xleftmin));
        yleft = yleftmin + floor(rng.rand() * (yleftmax - yleftmin));
        break;
      }
      if (++tries >= P.max_tries) { failed = true; break; }
    }
    const double th[3] = {-wtheta, wtheta, 180.0 - wtheta};
    px[0] = xleft; py[0] = yleft;
    for (int k = 0; k < 3; ++k) {
      const double a = kPiS * (th[k] / 180.0);   // theta2dir (geometry.py:177-188)
      double dx = cos(a), dy = sin(a);
      s_unit(dx, dy);
      px[k + 1] = px[k] + hlen * dx; py[k + 1] = py[k] + hlen * dy;
    }
  }
  SLine cage[4];
  T *wall = reinterpret_cast<T *>(S.wall) + (long long)w * n_walls * 8;
  for (int k = 0; k < 4; ++k) {
    const int k1 = (k + 1) & 3;
    cage[k] = s_line(px[k], py[k], px[k1], py[k1]);
    if (k < n_walls) {
      // get_box_coords (primitives.py:177-195)
      double nx = -cage[k].a, ny = -cage[k].b;
      s_unit(nx, ny);
      double dx = px[k1] - px[k], dy = py[k1] - py[k];
      const double dist = sqrt(dx * dx + dy * dy);
      s_unit(dx, dy);
      const double x1 = px[k], y1 = py[k];
      const double x2 = x1 + dist * dx, y2 = y1 + dist * dy;
      const double x3 = x2 - Tk * nx, y3 = y2 - Tk * ny;
      const double x4 = x3 - dist * dx, y4 = y3 - dist * dy;
      T *o = wall + k * 8;
      o[0] = (T)x1; o[1] = (T)y1; o[2] = (T)x2; o[3] = (T)y2; o[4] = (T)x3; o[5] = (T)y3; o[6] = (T)x4; o[7] = (T)y4;
    }
  }
  for (int k = 4; k < n_walls; ++k)
    for (int j = 0; j < 8; ++j) wall[k * 8 + j] = T(0);
  // ---- balls: add_balls (dataio.py:328-379) ----
  double bx[PPB_MAX_BALLS], by[PPB_MAX_BALLS], br[PPB_MAX_BALLS];
  for (int b = 0; b < n_balls && !failed; ++b) {
    int outer = 0;
    for (;;) {
      const double r = floor(P.mn_ball + rng.rand() * (P.mx_ball - P.mn_ball));
      double x = 0, y = 0;
      int tries = 0;
      for (;;) {   // find_point_within_lines(r + T + 2)
        x = rint(px[0] + rng.rand() * (px[2] - px[0]));
        y = rint(py[1] + rng.rand() * (py[3] - py[1]));
        bool ok = true;
        for (int k = 0; k < 4; ++k) {
          const double d = (cage[k].a * x + cage[k].b * y + cage[k].c) / sqrt(cage[k].a * cage[k].a + cage[k].b * cage[k].b);
          if (d <= r + Tk + 2.0) ok = false;
        }
        if (ok) break;
        if (++tries >= P.max_tries) { failed = true; break; }
      }
      if (failed) break;
      bool isok = true;
      for (int j = 0; j < b; ++j) {
        const double ddx = x - bx[j], ddy = y - by[j];
        isok = isok && (sqrt(ddx * ddx + ddy * ddy) > br[j] + r);
      }
      if (isok) { bx[b] = x; by[b] = y; br[b] = r; break; }
      if (++outer >= P.max_tries) { failed = true; break; }
    }
  }
  // ---- state + initial kick: apply_force (dataio.py:399-412), vel = 0.5 T^2 F / m (primitives.py:580-594) ----
  const long long base = (long long)w * n_balls;
  const double kPi = 3.141592653589793;
  for (int b = 0; b < n_balls; ++b) {
    T2 p, v, q;
    T4 z;
    z.x = z.y = z.z = z.w = T(0);
    if (failed) {
      p.x = p.y = v.x = v.y = T(0);
      q.x = T(-1); q.y = T(1);
    } else {
      const double mass = (4 / 3.0) * kPi * pow(br[b] / 10.0, 3.0);
      const double rnd1 = rng.rand();
      (void)rng.rand();   // rnd2 is drawn and never used (dataio.py:399)
      const double fmag = P.mn_force + floor(rnd1 * (P.mx_force - P.mn_force));
      double th = rnd1 * kPi;
      if (rng.rand() > 0.5) th = -th;
      const double fx = fmag * cos(th), fy = fmag * sin(th);
      const double s = 0.5 * P.force_t * P.force_t, im = 1.0 / mass;
      p.x = (T)bx[b]; p.y = (T)by[b];
      v.x = (T)(s * (fx * im)); v.y = (T)(s * (fy * im));
      q.x = (T)br[b]; q.y = (T)mass;
    }
    reinterpret_cast<T2 *>(S.pos)[base + b] = p;
    reinterpret_cast<T2 *>(S.vel)[base + b] = v;
    reinterpret_cast<T2 *>(S.rm)[base + b] = q;
    reinterpret_cast<T *>(S.tcol)[base + b] = T(0);
    S.partner[base + b] = -1;
    st4(reinterpret_cast<T4 *>(S.aux) + base + b, z);
  }
  // Dynamics._include_object: every ball queued (primitives.py:556-563)
  Hdr<T> h;
  memset(&h, 0, sizeof h);
  h.tmin = T(0);
  h.step = (uint32_t)step_value;
  h.flags = PPB_FLAG_PENDING;
  reinterpret_cast<Hdr<T> *>(S.hdr)[w] = h;
  S.nev[w] = 0u;
  if (failed) atomicAdd(fail_count, 1);
}

}  // namespace ppb
