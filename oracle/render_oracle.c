/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of World.generate_image()
 * (reference primitives.py:33-77, 129-174, 237-257, 306-317, 415-445, 876-877,
 * 991-1008).  Never used on the product path.
 *
 * PARITY UNPINNED: the reference draws through pycairo/libcairo, which is not part
 * of the reference tree, has no version stated there, is not installed in this
 * image, and the reference ships no golden images.  This file restates the Cairo
 * calls' drawing semantics:
 *   - (H, W, 4) uint8 canvas, all 255; array channel order [Color.r, Color.g,
 *     Color.b, alpha] (primitives.py:52-57, 73, 164-167 swizzle the colours so that
 *     ARGB32's little-endian BGRA byte order reads that way);
 *   - every object, in insertion order, is composited with OVER on premultiplied
 *     8-bit pixels using pixman's rounded multiply;
 *   - antialiasing: the 8-bit mask of a shape is floor(255 * exact_area + 0.5) where
 *     exact_area is the area of (shape ∩ pixel).  Cairo itself samples coverage on a
 *     15-row sub-pixel grid and flattens arcs to polylines with 0.1 px tolerance, so
 *     its edge pixels can differ from exact-area coverage by a few percent of full
 *     scale; interior and exterior pixels are identical.  Stated tolerance against
 *     real Cairo output: |diff| <= 26/255 on edge pixels, 0 elsewhere (an estimate,
 *     to be validated on a box that has pycairo);
 *   - walls: solid fColor through the mask of the wall quad, restricted to the
 *     integer rectangle the reference pastes its pre-rendered sprite into
 *     (primitives.py:242-245 GenericWall, :310-312 Wall);
 *   - balls: the pre-rendered sprite of get_ball_im (ring of sColor, then disc of
 *     fColor), pasted with its centre at the integer-rounded position (Python 2
 *     round(): half away from zero, geometry.py:25-29; primitives.py:428) and clipped
 *     to the canvas (primitives.py:429-440).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t mul_un8(uint32_t a, uint32_t b) {
  uint32_t t = a * b + 0x80u;
  return ((t >> 8) + t) >> 8;
}
/* premultiplied pixel as 4 bytes r,g,b,a */
typedef struct { uint8_t c[4]; } px_t;

static inline px_t px_over(px_t src, px_t dst) {
  px_t o;
  uint32_t ia = 255u - src.c[3];
  for (int k = 0; k < 4; ++k) {
    uint32_t v = src.c[k] + mul_un8(dst.c[k], ia);
    o.c[k] = (uint8_t)(v > 255u ? 255u : v);
  }
  return o;
}
static inline px_t px_in(px_t col, uint32_t m) {
  px_t o;
  for (int k = 0; k < 4; ++k) o.c[k] = (uint8_t)mul_un8(col.c[k], m);
  return o;
}
/* cairo: double -> 16 bit (x*65535+0.5), pixman: 16 -> 8 bit by >> 8 */
static inline uint8_t col8(double v) {
  if (v < 0) v = 0;
  if (v > 1) v = 1;
  return (uint8_t)(((uint32_t)(v * 65535.0 + 0.5)) >> 8);
}
static px_t premul(const float *rgba) {
  px_t p;
  double a = rgba[3];
  p.c[0] = col8(rgba[0] * a); p.c[1] = col8(rgba[1] * a); p.c[2] = col8(rgba[2] * a); p.c[3] = col8(a);
  return p;
}

/* ---- exact area of (disc radius R at origin) ∩ [x0,x1]x[y0,y1], integrating over y ---- */
static double seg_int(double y, double R) { /* antiderivative of sqrt(R^2 - y^2) */
  double q = R * R - y * y;
  if (q < 0) q = 0;
  double s = y / R;
  if (s > 1) s = 1;
  if (s < -1) s = -1;
  return 0.5 * (y * sqrt(q) + R * R * asin(s));
}
static int cmp_d(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
double oracle_disc_box_area(double R, double x0, double x1, double y0, double y1) {
  if (!(R > 0)) return 0.0;
  double lo = y0 > -R ? y0 : -R, hi = y1 < R ? y1 : R;
  if (!(lo < hi)) return 0.0;
  double br[8];
  int n = 0;
  br[n++] = lo;
  br[n++] = hi;
  double xs[2] = {x0, x1};
  for (int i = 0; i < 2; ++i) {
    double q = R * R - xs[i] * xs[i];
    if (q > 0) {
      double s = sqrt(q);
      if (-s > lo && -s < hi) br[n++] = -s;
      if (s > lo && s < hi) br[n++] = s;
    }
  }
  qsort(br, n, sizeof(double), cmp_d);
  double area = 0;
  for (int i = 0; i + 1 < n; ++i) {
    double u = br[i], v = br[i + 1];
    if (!(v > u)) continue;
    double ym = 0.5 * (u + v);
    double q = R * R - ym * ym;
    double w = sqrt(q > 0 ? q : 0);
    double right_is_arc = w < x1, left_is_arc = -w > x0;
    double r = right_is_arc ? w : x1, l = left_is_arc ? -w : x0;
    if (!(r > l)) continue;
    double I = seg_int(v, R) - seg_int(u, R);
    double part = (right_is_arc ? I : x1 * (v - u)) - (left_is_arc ? -I : x0 * (v - u));
    area += part;
  }
  return area;
}

/* ---- exact area of (convex polygon ∩ unit pixel at (px,py)) ---- */
double oracle_quad_pixel_area(const double *q /* 4 x (x,y) */, double px, double py) {
  double ax[16], ay[16], bx[16], by[16];
  int n = 4;
  for (int i = 0; i < 4; ++i) { ax[i] = q[2 * i] - px; ay[i] = q[2 * i + 1] - py; }
  for (int pl = 0; pl < 4; ++pl) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
      int j = (i + 1) % n;
      double ci, cj;
      switch (pl) {
        case 0: ci = ax[i]; cj = ax[j]; break;
        case 1: ci = 1.0 - ax[i]; cj = 1.0 - ax[j]; break;
        case 2: ci = ay[i]; cj = ay[j]; break;
        default: ci = 1.0 - ay[i]; cj = 1.0 - ay[j]; break;
      }
      if (ci >= 0) { bx[m] = ax[i]; by[m] = ay[i]; ++m; }
      if ((ci >= 0) != (cj >= 0)) {
        double t = ci / (ci - cj);
        bx[m] = ax[i] + t * (ax[j] - ax[i]);
        by[m] = ay[i] + t * (ay[j] - ay[i]);
        ++m;
      }
    }
    n = m;
    memcpy(ax, bx, sizeof(double) * n);
    memcpy(ay, by, sizeof(double) * n);
    if (n == 0) return 0.0;
  }
  double s = 0;
  for (int i = 0; i < n; ++i) {
    int j = (i + 1) % n;
    s += ax[i] * ay[j] - ax[j] * ay[i];
  }
  s = 0.5 * fabs(s);
  return s > 1 ? 1 : s;
}

/* get_ball_im, primitives.py:33-61: (sz, sz) premultiplied pixels, sz = 2*(r+sT) */
void oracle_ball_sprite(int r, int s_thick, const float *fill, const float *stroke, uint8_t *out) {
  int sz = 2 * (r + s_thick);
  double c = r + s_thick;
  px_t f8 = premul(fill), s8 = premul(stroke);
  for (int py = 0; py < sz; ++py)
    for (int px = 0; px < sz; ++px) {
      double x0 = px - c, x1 = px + 1 - c, y0 = py - c, y1 = py + 1 - c;
      double ri = r - 0.5 * s_thick;
      double cs = oracle_disc_box_area(r + 0.5 * s_thick, x0, x1, y0, y1) -
                  oracle_disc_box_area(ri > 0 ? ri : 0, x0, x1, y0, y1);
      double cf = oracle_disc_box_area(r, x0, x1, y0, y1);
      if (cs < 0) cs = 0;
      if (cs > 1) cs = 1;
      if (cf < 0) cf = 0;
      if (cf > 1) cf = 1;
      uint32_t ms = (uint32_t)floor(255.0 * cs + 0.5), mf = (uint32_t)floor(255.0 * cf + 0.5);
      px_t p = {{0, 0, 0, 0}};                          /* paint(1,1,1,0): transparent, :45-46 */
      if (s_thick > 0) p = px_over(px_in(s8, ms), p);   /* stroke, :49-55 */
      p = px_over(px_in(f8, mf), p);                    /* fill, :57-59 */
      memcpy(out + ((size_t)py * sz + px) * 4, p.c, 4);
    }
}

static double py2_round(double x) { return round(x); } /* half away from zero */

/*
 * World.generate_image for n_worlds worlds.
 *   wall: [n_worlds][n_walls][4][2]; wall_kind[n_walls] (0 GenericWall, 1 Wall); wall_rgba [n_walls][4]
 *   pos: [n_worlds][n_balls][2]; rad: [n_worlds][n_balls] (<0 = empty, else integral)
 *   s_thick[n_balls], fill/stroke [n_balls][4]; order: insertion order (wall q -> q, ball b -> n_walls+b)
 *   out: [n_worlds][H][W][4]
 */
int oracle_render_frames(int n_worlds, int n_balls, int n_walls, const int32_t *order, const double *wall,
                         const int32_t *wall_kind, const float *wall_rgba, const double *pos, const double *rad,
                         const int32_t *s_thick, const float *fill, const float *stroke, int W, int H, uint8_t *out,
                         int n_threads) {
  /* pre-render the sprites that occur (one per ball slot and integer radius) */
  int rmin = 1 << 30, rmax = -1;
  for (size_t i = 0; i < (size_t)n_worlds * n_balls; ++i)
    if (rad[i] >= 0) {
      int r = (int)rad[i];
      if (r < rmin) rmin = r;
      if (r > rmax) rmax = r;
    }
  int nr = rmax >= rmin ? rmax - rmin + 1 : 0;
  uint8_t **spr = (uint8_t **)calloc((size_t)(nr > 0 ? nr : 1) * n_balls, sizeof(uint8_t *));
  for (size_t i = 0; i < (size_t)n_worlds * n_balls; ++i)
    if (rad[i] >= 0) {
      int b = (int)(i % n_balls), r = (int)rad[i];
      uint8_t **slot = &spr[(size_t)b * nr + (r - rmin)];
      if (!*slot) {
        int sz = 2 * (r + s_thick[b]);
        *slot = (uint8_t *)malloc((size_t)sz * sz * 4 + 4);
        oracle_ball_sprite(r, s_thick[b], fill + 4 * b, stroke + 4 * b, *slot);
      }
    }
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (int w = 0; w < n_worlds; ++w) {
    uint8_t *im = out + (size_t)w * W * H * 4;
    memset(im, 255, (size_t)W * H * 4);                                   /* primitives.py:876, 998-999 */
    for (int k = 0; k < n_walls + n_balls; ++k) {
      int q = order[k];
      if (q < n_walls) {
        const double *v = wall + ((size_t)w * n_walls + q) * 8;
        double x0 = v[0], x1 = v[0], y0 = v[1], y1 = v[1];
        for (int i = 1; i < 4; ++i) {
          if (v[2 * i] < x0) x0 = v[2 * i];
          if (v[2 * i] > x1) x1 = v[2 * i];
          if (v[2 * i + 1] < y0) y0 = v[2 * i + 1];
          if (v[2 * i + 1] > y1) y1 = v[2 * i + 1];
        }
        int bx0, bx1, by0, by1;
        if (wall_kind[q] == 1) {            /* Wall.imprint, :310-312 (float slice bounds truncate) */
          bx0 = (int)x0; by0 = (int)y0; bx1 = (int)x1; by1 = (int)y1;
        } else {                            /* GenericWall: :148-149, :226, :242-245 */
          bx0 = (int)py2_round(x0); by0 = (int)py2_round(y0);
          bx1 = bx0 + (int)ceil(x1 - x0); by1 = by0 + (int)ceil(y1 - y0);
        }
        px_t col = premul(wall_rgba + 4 * q);
        int px_lo = bx0 > 0 ? bx0 : 0, px_hi = bx1 < W ? bx1 : W;
        int py_lo = by0 > 0 ? by0 : 0, py_hi = by1 < H ? by1 : H;
        for (int py = py_lo; py < py_hi; ++py)
          for (int px = px_lo; px < px_hi; ++px) {
            double a = oracle_quad_pixel_area(v, px, py);
            uint32_t m = (uint32_t)floor(255.0 * a + 0.5);
            if (!m) continue;
            px_t d;
            memcpy(d.c, im + ((size_t)py * W + px) * 4, 4);
            d = px_over(px_in(col, m), d);
            memcpy(im + ((size_t)py * W + px) * 4, d.c, 4);
          }
      } else {
        int b = q - n_walls;
        double rr = rad[(size_t)w * n_balls + b];
        if (rr < 0) continue;
        int r = (int)rr, off = r + s_thick[b], sz = 2 * off;              /* :420 */
        const uint8_t *s = spr[(size_t)b * nr + (r - rmin)];
        int X = (int)py2_round(pos[((size_t)w * n_balls + b) * 2]) - off;     /* :428 */
        int Y = (int)py2_round(pos[((size_t)w * n_balls + b) * 2 + 1]) - off;
        for (int sy = 0; sy < sz; ++sy) {
          int py = Y + sy;
          if (py < 0 || py >= H) continue;
          for (int sx = 0; sx < sz; ++sx) {
            int px = X + sx;
            if (px < 0 || px >= W) continue;
            px_t sp, d;
            memcpy(sp.c, s + ((size_t)sy * sz + sx) * 4, 4);
            if (!sp.c[3]) continue;
            memcpy(d.c, im + ((size_t)py * W + px) * 4, 4);
            d = px_over(sp, d);
            memcpy(im + ((size_t)py * W + px) * 4, d.c, 4);
          }
        }
      }
    }
  }
  for (size_t i = 0; i < (size_t)(nr > 0 ? nr : 1) * n_balls; ++i) free(spr[i]);
  free(spr);
  return 0;
}
