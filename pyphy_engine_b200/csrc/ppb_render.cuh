// ppb_render.cuh -- K3: rasterise World.generate_image() for every world (sm_100a).
//
// Drawing model (restatement of the reference's Cairo calls, primitives.py:33-77, 129-174,
// 237-257, 306-317, 415-445, 876-877, 991-1008; see DESIGN.md "Frame model"):
//   * canvas (H, W, 4) uint8, all 255; channel order of the returned array is
//     [Color.r, Color.g, Color.b, alpha] (the reference swizzles colours so that ARGB32's
//     BGRA byte order comes out that way);
//   * objects are composited in insertion order with Cairo's OVER operator on premultiplied
//     8-bit pixels, using pixman's rounded multiply  MUL(a,b) = ((a*b+128) + ((a*b+128)>>8)) >> 8;
//   * wall: solid fColor through an antialiased mask = exact area of (wall quad ∩ pixel),
//     clipped to the integer-placed sprite rectangle the reference pastes (primitives.py:242-245,
//     310-312); 8-bit mask m = floor(255*area + 0.5);
//   * ball: pre-rendered sprite of side 2*(r+sThick) (ring of sColor between r-sThick/2 and
//     r+sThick/2, then the fColor disc of radius r, both antialiased with exact pixel-area
//     coverage), pasted with its centre at the integer-rounded ball position (round half away
//     from zero, geometry.py:25-29, primitives.py:428) and clipped to the canvas.
//
// One CTA rasterises one 64x64 tile of one world's canvas; each thread owns runs of 4 pixels
// (one 128-bit store), the world's walls and balls are staged in shared memory.
#pragma once
#include "ppb_common.cuh"

namespace ppb {

__device__ __forceinline__ uint32_t mul_un8(uint32_t a, uint32_t b) {
  uint32_t t = a * b + 0x80u;
  return ((t >> 8) + t) >> 8;
}
// dst = src OVER dst, premultiplied, bytes r|g<<8|b<<16|a<<24
__device__ __forceinline__ uint32_t over_px(uint32_t src, uint32_t dst) {
  const uint32_t ia = 255u - (src >> 24);
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t s = (src >> (8 * c)) & 0xffu, d = (dst >> (8 * c)) & 0xffu;
    uint32_t v = s + mul_un8(d, ia);
    v = v > 255u ? 255u : v;
    out |= v << (8 * c);
  }
  return out;
}
// (solid colour IN mask m)
__device__ __forceinline__ uint32_t in_mask(uint32_t col, uint32_t m) {
  uint32_t out = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) out |= mul_un8((col >> (8 * c)) & 0xffu, m) << (8 * c);
  return out;
}

// colour component -> 8 bit the way cairo/pixman do: 16-bit (x*65535+0.5) then >> 8
__host__ __device__ inline uint32_t color_to_u8(double v) {
  if (v < 0) v = 0;
  if (v > 1) v = 1;
  return ((uint32_t)(v * 65535.0 + 0.5)) >> 8;
}
__host__ __device__ inline uint32_t premul_rgba8(const float c[4]) {
  const double a = c[3];
  return color_to_u8(c[0] * a) | (color_to_u8(c[1] * a) << 8) | (color_to_u8(c[2] * a) << 16) |
         (color_to_u8(a) << 24);
}

// ---- exact area of (disc of radius R centred at the origin) ∩ [x0,x1]x[y0,y1] ----
__device__ inline double circ_H(double x, double R) {  // antiderivative of sqrt(R^2 - x^2)
  double q = R * R - x * x;
  q = q > 0 ? q : 0;
  double s = x / R;
  s = s > 1 ? 1 : (s < -1 ? -1 : s);
  return 0.5 * (x * sqrt(q) + R * R * asin(s));
}
__device__ inline double disc_box_area(double R, double x0, double x1, double y0, double y1) {
  if (!(R > 0)) return 0.0;
  // clip the x range to the disc
  double a = x0 > -R ? x0 : -R, b = x1 < R ? x1 : R;
  if (!(a < b)) return 0.0;
  // breakpoints where the chord end +-h(x) crosses y0 or y1
  double bp[6];
  int n = 0;
  bp[n++] = a;
  const double ys[2] = {y0, y1};
  double cand[4];
  int nc = 0;
  for (int i = 0; i < 2; ++i) {
    double q = R * R - ys[i] * ys[i];
    if (q > 0) {
      double s = sqrt(q);
      cand[nc++] = -s;
      cand[nc++] = s;
    }
  }
  // insertion sort of the candidates inside (a, b)
  for (int i = 0; i < nc; ++i)
    for (int j = i + 1; j < nc; ++j)
      if (cand[j] < cand[i]) {
        double t = cand[i];
        cand[i] = cand[j];
        cand[j] = t;
      }
  for (int i = 0; i < nc; ++i)
    if (cand[i] > bp[n - 1] && cand[i] < b) bp[n++] = cand[i];
  bp[n++] = b;
  double area = 0.0;
  for (int i = 0; i + 1 < n; ++i) {
    const double u = bp[i], v = bp[i + 1];
    const double xm = 0.5 * (u + v);
    double q = R * R - xm * xm;
    const double h = sqrt(q > 0 ? q : 0);
    const double top = h < y1 ? h : y1;      // min(y1, h)
    const double bot = -h > y0 ? -h : y0;    // max(y0, -h)
    if (!(top > bot)) continue;
    // integrand = (h<y1 ? h : y1) - (-h>y0 ? -h : y0), constant structure on [u, v]
    double part = 0.0;
    const double Hh = circ_H(v, R) - circ_H(u, R);
    part += (h < y1) ? Hh : y1 * (v - u);
    part -= (-h > y0) ? -Hh : y0 * (v - u);
    area += part;
  }
  return area;
}

// get_ball_im (primitives.py:33-61): one thread per sprite pixel
static __global__ void build_ball_sprites_kernel(uint32_t *atlas, RenderLayout RL, int n_balls, const uint32_t *fill8,
                                          const uint32_t *stroke8) {
  const int slot = blockIdx.y;
  const int ri = blockIdx.z;  // radius index
  if (slot >= n_balls) return;
  const int r = RL.r_min + ri;
  const int sT = RL.s_thick[slot];
  const int sz = 2 * (r + sT);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= sz * sz) return;
  const int px = p % sz, py = p / sz;
  const double c = (double)(r + sT);
  const double x0 = px - c, x1 = px + 1 - c, y0 = py - c, y1 = py + 1 - c;
  const double Ro = r + 0.5 * sT, Ri = r - 0.5 * sT;
  double cs = disc_box_area(Ro, x0, x1, y0, y1) - disc_box_area(Ri > 0 ? Ri : 0, x0, x1, y0, y1);
  double cf = disc_box_area((double)r, x0, x1, y0, y1);
  cs = cs < 0 ? 0 : (cs > 1 ? 1 : cs);
  cf = cf < 0 ? 0 : (cf > 1 ? 1 : cf);
  const uint32_t ms = (uint32_t)floor(255.0 * cs + 0.5), mf = (uint32_t)floor(255.0 * cf + 0.5);
  uint32_t px8 = 0u;                                  // transparent (primitives.py:45-46)
  if (sT > 0) px8 = over_px(in_mask(stroke8[slot], ms), px8);  // stroke (primitives.py:49-55)
  px8 = over_px(in_mask(fill8[slot], mf), px8);       // fill   (primitives.py:57-59)
  uint8_t *basep = reinterpret_cast<uint8_t *>(atlas) + (size_t)slot * RL.sprite_slot_stride +
                   (size_t)ri * RL.sprite_stride;
  reinterpret_cast<uint32_t *>(basep)[p] = px8;
}

// ---- exact area of (convex quad ∩ pixel [px,px+1]x[py,py+1]) by Sutherland-Hodgman ----
__device__ inline float quad_pixel_area(const float *qv /*8*/, float px, float py) {
  float ax[10], ay[10], bx[10], by[10];
  int n = 4;
  for (int i = 0; i < 4; ++i) {
    ax[i] = qv[2 * i] - px;
    ay[i] = qv[2 * i + 1] - py;
  }
  // clip against x>=0, x<=1, y>=0, y<=1
  for (int pl = 0; pl < 4; ++pl) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const int j = (i + 1 == n) ? 0 : i + 1;
      float ci, cj;
      if (pl == 0) { ci = ax[i]; cj = ax[j]; }
      else if (pl == 1) { ci = 1.f - ax[i]; cj = 1.f - ax[j]; }
      else if (pl == 2) { ci = ay[i]; cj = ay[j]; }
      else { ci = 1.f - ay[i]; cj = 1.f - ay[j]; }
      if (ci >= 0.f) {
        bx[m] = ax[i]; by[m] = ay[i]; ++m;
      }
      if ((ci >= 0.f) != (cj >= 0.f)) {
        const float t = ci / (ci - cj);
        bx[m] = ax[i] + t * (ax[j] - ax[i]);
        by[m] = ay[i] + t * (ay[j] - ay[i]);
        ++m;
      }
    }
    n = m;
    for (int i = 0; i < n; ++i) { ax[i] = bx[i]; ay[i] = by[i]; }
    if (n == 0) return 0.f;
  }
  float s = 0.f;
  for (int i = 0; i < n; ++i) {
    const int j = (i + 1 == n) ? 0 : i + 1;
    s += ax[i] * ay[j] - ax[j] * ay[i];
  }
  s = 0.5f * fabsf(s);
  return s > 1.f ? 1.f : s;
}

struct SWall {
  float v[8];        // quad vertices
  float x0, x1, y0, y1;   // bounding box of the quad
  int bx0, bx1, by0, by1; // integer rectangle the reference pastes the sprite into
  int axis;          // 1 = axis-aligned rectangle (coverage separable)
  uint32_t col;
};
struct SBall {
  int x0, y0;        // top-left of the sprite on the canvas
  int sz;            // sprite side, 0 = empty slot
  const uint32_t *tex;
};

#define PPB_TILE 64

template <typename T>
__global__ void __launch_bounds__(256) render_kernel(StatePtrs S, Layout L, RenderLayout RL, const uint32_t *atlas,
                                                     uint8_t *frames) {
  typedef typename Vec2<T>::type T2;
  __shared__ SWall sw[PPB_MAX_WALLS];
  __shared__ SBall sb[PPB_MAX_BALLS];
  const int tiles_x = (RL.W + PPB_TILE - 1) / PPB_TILE;
  const int tiles_y = (RL.H + PPB_TILE - 1) / PPB_TILE;
  const int tiles = tiles_x * tiles_y;
  const int w = blockIdx.x / tiles;
  const int tile = blockIdx.x - w * tiles;
  const int tx0 = (tile % tiles_x) * PPB_TILE, ty0 = (tile / tiles_x) * PPB_TILE;
  const int nB = L.n_balls, nWl = L.n_walls;

  if (threadIdx.x < nWl) {
    const int q = threadIdx.x;
    const T *wv = reinterpret_cast<const T *>(S.wall) + ((long long)w * nWl + q) * 8;
    SWall s;
    float x0 = 1e30f, x1 = -1e30f, y0 = 1e30f, y1 = -1e30f;
    double dx0 = 1e300, dy0 = 1e300, dx1 = -1e300, dy1 = -1e300;
    for (int i = 0; i < 4; ++i) {
      const double vx = (double)wv[2 * i], vy = (double)wv[2 * i + 1];
      s.v[2 * i] = (float)vx;
      s.v[2 * i + 1] = (float)vy;
      x0 = fminf(x0, (float)vx); x1 = fmaxf(x1, (float)vx);
      y0 = fminf(y0, (float)vy); y1 = fmaxf(y1, (float)vy);
      dx0 = fmin(dx0, vx); dx1 = fmax(dx1, vx);
      dy0 = fmin(dy0, vy); dy1 = fmax(dy1, vy);
    }
    s.x0 = x0; s.x1 = x1; s.y0 = y0; s.y1 = y1;
    // axis-aligned when every edge is horizontal or vertical
    int ax = 1;
    for (int i = 0; i < 4; ++i) {
      const int j = (i + 1) & 3;
      if (s.v[2 * i] != s.v[2 * j] && s.v[2 * i + 1] != s.v[2 * j + 1]) ax = 0;
    }
    s.axis = ax;
    if (RL.wall[q].kind == 1) {  // Wall.imprint: srcIm[y : y+sz.y, x : x+sz.x] (primitives.py:310-312)
      s.bx0 = (int)dx0; s.by0 = (int)dy0;
      s.bx1 = (int)dx1; s.by1 = (int)dy1;
    } else {                     // GenericWall: pos_ rounded, sprite ceil-sized (primitives.py:148-149, 242-245)
      s.bx0 = (int)round(dx0); s.by0 = (int)round(dy0);
      s.bx1 = s.bx0 + (int)ceil(dx1 - dx0);
      s.by1 = s.by0 + (int)ceil(dy1 - dy0);
    }
    s.col = RL.wall[q].rgba8;
    sw[q] = s;
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + nB) {
    const int b = threadIdx.x - 32;
    const T2 p = reinterpret_cast<const T2 *>(S.pos)[(long long)w * nB + b];
    const T rad = reinterpret_cast<const T2 *>(S.rm)[(long long)w * nB + b].x;
    SBall s;
    s.sz = 0; s.x0 = 0; s.y0 = 0; s.tex = atlas;
    const int r = (int)rad;
    if (rad >= T(0) && r >= RL.r_min && r <= RL.r_max) {
      const int off = r + RL.s_thick[b];   // floor(sz/2), primitives.py:420
      s.sz = 2 * off;
      s.x0 = (int)round((double)p.x) - off;  // x_asint: round half away from zero (geometry.py:25-29)
      s.y0 = (int)round((double)p.y) - off;
      s.tex = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(atlas) +
                                                 (size_t)b * RL.sprite_slot_stride +
                                                 (size_t)(r - RL.r_min) * RL.sprite_stride);
    }
    sb[b] = s;
  }
  __syncthreads();

  uint8_t *out = frames + (size_t)w * RL.W * RL.H * 4;
  const bool vec_ok = (RL.W & 3) == 0;
  // 64x64 tile = 64 rows x 16 groups of 4 px; 256 threads -> 4 groups each
  for (int it = 0; it < 4; ++it) {
    const int gidx = it * 256 + threadIdx.x;
    const int gy = ty0 + (gidx >> 4);
    const int gx = tx0 + ((gidx & 15) << 2);
    if (gy >= RL.H || gx >= RL.W) continue;
    uint32_t c[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    for (int k = 0; k < L.n_obj; ++k) {
      const int q = L.order[k];
      if (q < nWl) {
        const SWall &s = sw[q];
        if ((float)(gy + 1) <= s.y0 || (float)gy >= s.y1 || (float)(gx + 4) <= s.x0 || (float)gx >= s.x1) continue;
        if (gy < s.by0 || gy >= s.by1) continue;
        if (s.axis) {
          const float oy = fminf((float)(gy + 1), s.y1) - fmaxf((float)gy, s.y0);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const int x = gx + p;
            if (x < s.bx0 || x >= s.bx1) continue;
            const float ox = fminf((float)(x + 1), s.x1) - fmaxf((float)x, s.x0);
            if (ox <= 0.f) continue;
            const uint32_t m = (uint32_t)floorf(255.f * ox * oy + 0.5f);
            if (m) c[p] = over_px(m == 255u ? s.col : in_mask(s.col, m), c[p]);
          }
        } else {
#pragma unroll 1
          for (int p = 0; p < 4; ++p) {
            const int x = gx + p;
            if (x < s.bx0 || x >= s.bx1) continue;
            const float a = quad_pixel_area(s.v, (float)x, (float)gy);
            const uint32_t m = (uint32_t)floorf(255.f * a + 0.5f);
            if (m) c[p] = over_px(m == 255u ? s.col : in_mask(s.col, m), c[p]);
          }
        }
      } else {
        const SBall &s = sb[q - nWl];
        const int dy = gy - s.y0;
        const int dx0 = gx - s.x0;
        if (dy < 0 || dy >= s.sz || dx0 + 3 < 0 || dx0 >= s.sz) continue;
        const uint32_t *row = s.tex + dy * s.sz;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int dx = dx0 + p;
          if (dx < 0 || dx >= s.sz) continue;
          const uint32_t t = __ldg(row + dx);
          if (t >> 24) c[p] = over_px(t, c[p]);
        }
      }
    }
    uint8_t *dst = out + ((size_t)gy * RL.W + gx) * 4;
    if (vec_ok) {
      *reinterpret_cast<uint4 *>(dst) = make_uint4(c[0], c[1], c[2], c[3]);
    } else {
      for (int p = 0; p < 4 && gx + p < RL.W; ++p) reinterpret_cast<uint32_t *>(dst)[p] = c[p];
    }
  }
}

}  // namespace ppb
