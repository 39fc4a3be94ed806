/*
 * portable_trig.h -- sin / asin / acos in IEEE fp64 built only from + - * /
 * and sqrt, with no fused multiply-add and no table look-ups.
 *
 * Why: the reference computes the ball-ball time of impact with
 * math.acos / math.asin / np.sin (geometry.py:174, 443-448).  CUDA's libdevice
 * and glibc differ in the last bit of those functions, which would make fp64
 * trajectories on the GPU drift away from a CPU restatement after the first
 * ball-ball contact.  The fp64 ("compat") kernels therefore use these
 * implementations, which give the same bits on the host (gcc,
 * -ffp-contract=off) and on the device (nvcc, -fmad=false), so a CPU checker
 * that includes this header can be compared with the GPU bit for bit.
 * Accuracy is < 1 ulp (classic Sun fdlibm polynomial kernels, written out
 * from the published algorithm; argument range limited to what the
 * billiards path needs: |x| <= ~2^20 for sin, |x| <= 1 for asin / acos).
 *
 * C and CUDA C++ compatible.
 */
#ifndef PYPHY_B200_PORTABLE_TRIG_H
#define PYPHY_B200_PORTABLE_TRIG_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PTRIG_FN static __host__ __device__ __forceinline__
#else
#define PTRIG_FN static inline
#endif

PTRIG_FN double ptrig_from_words(uint32_t hi, uint32_t lo) {
  uint64_t u = ((uint64_t)hi << 32) | (uint64_t)lo;
  double d;
#if defined(__CUDA_ARCH__)
  d = __longlong_as_double((long long)u);
#else
  memcpy(&d, &u, sizeof d);
#endif
  return d;
}
PTRIG_FN uint32_t ptrig_hi(double d) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2hiint(d);
#else
  uint64_t u;
  memcpy(&u, &d, sizeof u);
  return (uint32_t)(u >> 32);
#endif
}
PTRIG_FN double ptrig_clear_lo(double d) { return ptrig_from_words(ptrig_hi(d), 0u); }

/* sin on [-pi/4, pi/4] with tail y */
PTRIG_FN double ptrig_ksin(double x, double y, int iy) {
  const double S1 = ptrig_from_words(0xBFC55555u, 0x55555549u);
  const double S2 = ptrig_from_words(0x3F811111u, 0x1110F8A6u);
  const double S3 = ptrig_from_words(0xBF2A01A0u, 0x19C161D5u);
  const double S4 = ptrig_from_words(0x3EC71DE3u, 0x57B1FE7Du);
  const double S5 = ptrig_from_words(0xBE5AE5E6u, 0x8A2B9CEBu);
  const double S6 = ptrig_from_words(0x3DE5D93Au, 0x5ACFD57Cu);
  uint32_t ix = ptrig_hi(x) & 0x7fffffffu;
  if (ix < 0x3e400000u) { if ((int)x == 0) return x; }
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (iy == 0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

/* cos on [-pi/4, pi/4] with tail y */
PTRIG_FN double ptrig_kcos(double x, double y) {
  const double C1 = ptrig_from_words(0x3FA55555u, 0x5555554Cu);
  const double C2 = ptrig_from_words(0xBF56C16Cu, 0x16C15177u);
  const double C3 = ptrig_from_words(0x3EFA01A0u, 0x19CB1590u);
  const double C4 = ptrig_from_words(0xBE927E4Fu, 0x809C52ADu);
  const double C5 = ptrig_from_words(0x3E21EE9Eu, 0xBDB4B1C4u);
  const double C6 = ptrig_from_words(0xBDA8FAE9u, 0xBE8838D4u);
  uint32_t ix = ptrig_hi(x) & 0x7fffffffu;
  if (ix < 0x3e400000u) { if ((int)x == 0) return 1.0; }
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  if (ix < 0x3FD33333u) return 1.0 - (0.5 * z - (z * r - x * y));
  double qx;
  if (ix > 0x3fe90000u) qx = 0.28125;
  else qx = ptrig_from_words(ix - 0x00200000u, 0u);
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

/* x = n*pi/2 + (y0 + y1), |y0| <= pi/4; valid for |x| < 2^20*pi/2 */
PTRIG_FN int ptrig_rem_pio2(double x, double *y0, double *y1) {
  const double invpio2 = ptrig_from_words(0x3FE45F30u, 0x6DC9C883u);
  const double pio2_1 = ptrig_from_words(0x3FF921FBu, 0x54400000u);
  const double pio2_1t = ptrig_from_words(0x3DD0B461u, 0x1A626331u);
  const double pio2_2 = ptrig_from_words(0x3DD0B461u, 0x1A600000u);
  const double pio2_2t = ptrig_from_words(0x3BA3198Au, 0x2E037073u);
  const double pio2_3 = ptrig_from_words(0x3BA3198Au, 0x2E000000u);
  const double pio2_3t = ptrig_from_words(0x397B839Au, 0x252049C1u);
  double ax = fabs(x);
  int n = (int)(ax * invpio2 + 0.5);
  double fn = (double)n;
  double r = ax - fn * pio2_1;
  double w = fn * pio2_1t;
  int j = (int)((ptrig_hi(ax) & 0x7fffffffu) >> 20);
  double a0 = r - w;
  int i = j - (int)((ptrig_hi(a0) >> 20) & 0x7ffu);
  if (i > 16) {
    double t = r;
    w = fn * pio2_2;
    r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    a0 = r - w;
    i = j - (int)((ptrig_hi(a0) >> 20) & 0x7ffu);
    if (i > 49) {
      t = r;
      w = fn * pio2_3;
      r = t - w;
      w = fn * pio2_3t - ((t - r) - w);
      a0 = r - w;
    }
  }
  double a1 = (r - a0) - w;
  if (x < 0) { *y0 = -a0; *y1 = -a1; return -n; }
  *y0 = a0; *y1 = a1;
  return n;
}

PTRIG_FN double ptrig_sin(double x) {
  uint32_t ix = ptrig_hi(x) & 0x7fffffffu;
  if (ix <= 0x3fe921fbu) return ptrig_ksin(x, 0.0, 0);
  if (ix >= 0x7ff00000u) return x - x;
  double y0, y1;
  int n = ptrig_rem_pio2(x, &y0, &y1);
  switch (n & 3) {
    case 0: return ptrig_ksin(y0, y1, 1);
    case 1: return ptrig_kcos(y0, y1);
    case 2: return -ptrig_ksin(y0, y1, 1);
    default: return -ptrig_kcos(y0, y1);
  }
}

PTRIG_FN double ptrig_asin_pq(double t, double *q) {
  const double pS0 = ptrig_from_words(0x3FC55555u, 0x55555555u);
  const double pS1 = ptrig_from_words(0xBFD4D612u, 0x03EB6F7Du);
  const double pS2 = ptrig_from_words(0x3FC9C155u, 0x0E884455u);
  const double pS3 = ptrig_from_words(0xBFA48228u, 0xB5688F3Bu);
  const double pS4 = ptrig_from_words(0x3F49EFE0u, 0x7501B288u);
  const double pS5 = ptrig_from_words(0x3F023DE1u, 0x0DFDF709u);
  const double qS1 = ptrig_from_words(0xC0033A27u, 0x1C8A2D4Bu);
  const double qS2 = ptrig_from_words(0x40002AE5u, 0x9C598AC8u);
  const double qS3 = ptrig_from_words(0xBFE6066Cu, 0x1B8D0159u);
  const double qS4 = ptrig_from_words(0x3FB3B8C5u, 0xB12E9282u);
  *q = 1.0 + t * (qS1 + t * (qS2 + t * (qS3 + t * qS4)));
  return t * (pS0 + t * (pS1 + t * (pS2 + t * (pS3 + t * (pS4 + t * pS5)))));
}

PTRIG_FN double ptrig_asin(double x) {
  const double pio2_hi = ptrig_from_words(0x3FF921FBu, 0x54442D18u);
  const double pio2_lo = ptrig_from_words(0x3C91A626u, 0x33145C07u);
  const double pio4_hi = ptrig_from_words(0x3FE921FBu, 0x54442D18u);
  uint32_t hx = ptrig_hi(x);
  uint32_t ix = hx & 0x7fffffffu;
  double p, q, w, t;
  if (ix >= 0x3ff00000u) {
    if (fabs(x) == 1.0) return x * pio2_hi + x * pio2_lo;
    return (x - x) / (x - x);
  }
  if (ix < 0x3fe00000u) {
    if (ix < 0x3e400000u) return x;
    t = x * x;
    p = ptrig_asin_pq(t, &q);
    w = p / q;
    return x + x * w;
  }
  w = 1.0 - fabs(x);
  t = w * 0.5;
  p = ptrig_asin_pq(t, &q);
  double s = sqrt(t);
  if (ix >= 0x3FEF3333u) {
    w = p / q;
    t = pio2_hi - (2.0 * (s + s * w) - pio2_lo);
  } else {
    w = ptrig_clear_lo(s);
    double c = (t - w * w) / (s + w);
    double r = p / q;
    p = 2.0 * s * r - (pio2_lo - 2.0 * c);
    q = pio4_hi - 2.0 * w;
    t = pio4_hi - (p - q);
  }
  return (hx >> 31) ? -t : t;
}

PTRIG_FN double ptrig_acos(double x) {
  const double pio2_hi = ptrig_from_words(0x3FF921FBu, 0x54442D18u);
  const double pio2_lo = ptrig_from_words(0x3C91A626u, 0x33145C07u);
  const double pi = ptrig_from_words(0x400921FBu, 0x54442D18u);
  uint32_t hx = ptrig_hi(x);
  uint32_t ix = hx & 0x7fffffffu;
  double z, p, q, r, w, s, c, df;
  if (ix >= 0x3ff00000u) {
    if (fabs(x) == 1.0) return (hx >> 31) ? pi + 2.0 * pio2_lo : 0.0;
    return (x - x) / (x - x);
  }
  if (ix < 0x3fe00000u) {
    if (ix <= 0x3c600000u) return pio2_hi + pio2_lo;
    z = x * x;
    p = ptrig_asin_pq(z, &q);
    r = p / q;
    return pio2_hi - (x - (pio2_lo - x * r));
  }
  if (hx >> 31) {
    z = (1.0 + x) * 0.5;
    p = ptrig_asin_pq(z, &q);
    s = sqrt(z);
    r = p / q;
    w = r * s - pio2_lo;
    return pi - 2.0 * (s + w);
  }
  z = (1.0 - x) * 0.5;
  s = sqrt(z);
  df = ptrig_clear_lo(s);
  c = (z - df * df) / (s + df);
  p = ptrig_asin_pq(z, &q);
  r = p / q;
  w = r * s + c;
  return 2.0 * (df + w);
}

#endif /* PYPHY_B200_PORTABLE_TRIG_H */
