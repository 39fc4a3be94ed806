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
//   * ball: pre-rendered sprite of side 2*(r+sThick) (ring of sColor of width sThick centred on radius r, then
//     the fColor disc of radius r, rasterised the way cairo's image backend does it -- csrc/ppb_cairo.cu),
//     pasted with its centre at the integer-rounded ball position (round half away
//     from zero, geometry.py:25-29, primitives.py:428) and clipped to the canvas.
//
// One CTA rasterises one 64x64 tile of one world's canvas; each thread owns runs of 4 pixels
// (one 128-bit store), the world's walls and balls are staged in shared memory.
#pragma once
#include <cuda.h>   // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint, see ppb_launch_impl.cuh)

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

// ball sprites (get_ball_im): csrc/ppb_cairo.cu

// ---- exact area of (convex quad ∩ pixel [px,px+1]x[py,py+1]) by Sutherland-Hodgman ----
static __device__ __noinline__ float quad_pixel_area(const float *qv /*8*/, float px, float py) {
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

// ---- packed 8-bit arithmetic (two channels per 32-bit multiply, as pixman does) ----
__device__ __forceinline__ uint32_t mul_un8x4(uint32_t x, uint32_t a) {
  uint32_t rb = (x & 0x00ff00ffu) * a + 0x00800080u;
  rb = ((rb + ((rb >> 8) & 0x00ff00ffu)) >> 8) & 0x00ff00ffu;
  uint32_t ag = ((x >> 8) & 0x00ff00ffu) * a + 0x00800080u;
  ag = (ag + ((ag >> 8) & 0x00ff00ffu)) & 0xff00ff00u;
  return rb | ag;
}
// src OVER dst for premultiplied pixels: every channel of the sum is <= 255 because
// src.c <= src.a and mul(dst.c, 255 - src.a) <= 255 - src.a, so a plain add cannot carry
__device__ __forceinline__ uint32_t over_fast(uint32_t src, uint32_t dst) {
  const uint32_t ia = 255u - (src >> 24);
  if (dst == 0xffffffffu) return src + ia * 0x01010101u;   // white, opaque canvas: mul(255, ia) = ia
  return src + mul_un8x4(dst, ia);
}

struct SWallGeom {
  float v[8];        // quad vertices
  float x0, x1, y0, y1;   // bounding box of the quad
  int bx0, bx1, by0, by1; // integer rectangle the reference pastes the sprite into
  int axis;          // 1 = axis-aligned rectangle (coverage separable)
};
struct SWall : SWallGeom {
  int fast;          // axis-aligned with integral y edges: the row pattern is constant between them
  int lo, hi;        // rows [lo, hi) can be touched by this wall
  uint32_t col;
};
struct SBall {
  int x0, y0;        // top-left of the sprite on the canvas
  int sz;            // sprite side, 0 = empty slot
  const uint32_t *tex;
};

#define PPB_TILE 64

template <typename T>
__device__ __forceinline__ void setup_wall(SWall &s, const T *wv, const RenderWall &rw) {
  float x0 = 1e30f, x1 = -1e30f, y0 = 1e30f, y1 = -1e30f;
  double dx0 = 1e300, dy0 = 1e300, dx1 = -1e300, dy1 = -1e300;
#pragma unroll
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
  int ax = 1;  // axis-aligned when every edge is horizontal or vertical
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    if (s.v[2 * i] != s.v[2 * j] && s.v[2 * i + 1] != s.v[2 * j + 1]) ax = 0;
  }
  s.axis = ax;
  s.fast = ax && floorf(y0) == y0 && floorf(y1) == y1;
  if (rw.kind == 1) {  // Wall.imprint: srcIm[y : y+sz.y, x : x+sz.x] (primitives.py:310-312)
    s.bx0 = (int)dx0; s.by0 = (int)dy0;
    s.bx1 = (int)dx1; s.by1 = (int)dy1;
  } else {             // GenericWall: pos_ rounded, sprite ceil-sized (primitives.py:148-149, 242-245)
    s.bx0 = (int)round(dx0); s.by0 = (int)round(dy0);
    s.bx1 = s.bx0 + (int)ceil(dx1 - dx0);
    s.by1 = s.by0 + (int)ceil(dy1 - dy0);
  }
  s.col = rw.rgba8;
  s.lo = max((int)floorf(y0), s.by0);
  s.hi = min((int)ceilf(y1), s.by1);
}

template <typename T>
__device__ __forceinline__ void setup_ball(SBall &s, const StatePtrs &S, const RenderLayout &RL, const uint32_t *atlas,
                                           int w, int nB, int b) {
  typedef typename Vec2<T>::type T2;
  const T2 p = reinterpret_cast<const T2 *>(S.pos)[(long long)w * nB + b];
  const T rad = reinterpret_cast<const T2 *>(S.rm)[(long long)w * nB + b].x;
  s.sz = 0; s.x0 = 0; s.y0 = 0; s.tex = atlas;
  const int r = (int)rad;
  if (rad >= T(0) && r >= RL.r_min && r <= RL.r_max) {
    const int off = r + RL.s_thick[b];     // floor(sz/2), primitives.py:420
    s.sz = 2 * off;
    s.x0 = (int)round((double)p.x) - off;  // x_asint: round half away from zero (geometry.py:25-29)
    s.y0 = (int)round((double)p.y) - off;
    s.tex = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(atlas) +
                                               (size_t)b * RL.sprite_slot_stride +
                                               (size_t)(r - RL.r_min) * RL.sprite_stride);
  }
}

// 8-bit mask of wall s on pixel (x, y)
__device__ __forceinline__ uint32_t wall_mask(const SWallGeom &s, int x, int y) {
  if (x < s.bx0 || x >= s.bx1 || y < s.by0 || y >= s.by1) return 0u;
  float a;
  if (s.axis) {
    const float ox = fminf((float)(x + 1), s.x1) - fmaxf((float)x, s.x0);
    const float oy = fminf((float)(y + 1), s.y1) - fmaxf((float)y, s.y0);
    if (ox <= 0.f || oy <= 0.f) return 0u;
    a = ox * oy;
  } else {
    if ((float)(x + 1) <= s.x0 || (float)x >= s.x1 || (float)(y + 1) <= s.y0 || (float)y >= s.y1) return 0u;
    a = quad_pixel_area(s.v, (float)x, (float)y);
  }
  return (uint32_t)floorf(255.f * a + 0.5f);
}
__device__ __forceinline__ uint32_t put_wall(const SWallGeom &s, uint32_t col, int x, int y, uint32_t c) {
  const uint32_t m = wall_mask(s, x, y);
  if (!m) return c;
  return over_fast(m == 255u ? col : mul_un8x4(col, m), c);
}
__device__ __forceinline__ uint32_t put_ball(const SBall &s, int x, int y, uint32_t c) {
  const int dy = y - s.y0, dx = x - s.x0;
  if ((unsigned)dy >= (unsigned)s.sz || (unsigned)dx >= (unsigned)s.sz) return c;
  const uint32_t t = __ldg(s.tex + dy * s.sz + dx);
  return (t >> 24) ? over_fast(t, c) : c;
}

// wall s composited onto the 4 pixels (x..x+3, y)
__device__ __forceinline__ void put_wall4(const SWallGeom &s, uint32_t col, int x, int y, uint32_t &c0, uint32_t &c1,
                                          uint32_t &c2, uint32_t &c3) {
  if (y < s.by0 || y >= s.by1 || x + 4 <= s.bx0 || x >= s.bx1) return;
  if (s.axis) {
    const float oy = fminf((float)(y + 1), s.y1) - fmaxf((float)y, s.y0);
    if (oy <= 0.f) return;
    if (oy == 1.f && (float)x >= s.x0 && (float)(x + 4) <= s.x1 && x >= s.bx0 && x + 4 <= s.bx1) {
      // all four pixels fully covered
      if ((col >> 24) == 255u) { c0 = c1 = c2 = c3 = col; }
      else { c0 = over_fast(col, c0); c1 = over_fast(col, c1); c2 = over_fast(col, c2); c3 = over_fast(col, c3); }
      return;
    }
  }
  c0 = put_wall(s, col, x, y, c0);
  c1 = put_wall(s, col, x + 1, y, c1);
  c2 = put_wall(s, col, x + 2, y, c2);
  c3 = put_wall(s, col, x + 3, y, c3);
}

// ------------------------------------------------------------------------------------
// general rasteriser: any insertion order, any canvas width; one CTA per 64x64 tile
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) render_general_kernel(StatePtrs S, Layout L, RenderLayout RL,
                                                             const uint32_t *atlas, uint8_t *frames) {
  __shared__ SWall sw[PPB_MAX_WALLS];
  __shared__ SBall sb[PPB_MAX_BALLS];
  const int tiles_x = (RL.W + PPB_TILE - 1) / PPB_TILE;
  const int tiles_y = (RL.H + PPB_TILE - 1) / PPB_TILE;
  const int tiles = tiles_x * tiles_y;
  const int w = blockIdx.x / tiles;
  const int tile = blockIdx.x - w * tiles;
  const int tx0 = (tile % tiles_x) * PPB_TILE, ty0 = (tile / tiles_x) * PPB_TILE;
  const int nB = L.n_balls, nWl = L.n_walls;
  if (threadIdx.x < nWl)
    setup_wall<T>(sw[threadIdx.x], reinterpret_cast<const T *>(S.wall) + ((long long)w * nWl + threadIdx.x) * 8,
                  RL.wall[threadIdx.x]);
  if (threadIdx.x >= 32 && threadIdx.x < 32 + nB) setup_ball<T>(sb[threadIdx.x - 32], S, RL, atlas, w, nB, threadIdx.x - 32);
  __syncthreads();
  uint32_t *out = reinterpret_cast<uint32_t *>(frames) + (size_t)w * RL.W * RL.H;
  for (int it = threadIdx.x; it < PPB_TILE * PPB_TILE; it += 256) {
    const int y = ty0 + (it >> 6), x = tx0 + (it & 63);
    if (y >= RL.H || x >= RL.W) continue;
    uint32_t c = 0xffffffffu;
    for (int k = 0; k < L.n_obj; ++k) {
      const int q = L.order[k];
      if (q < nWl) c = put_wall(sw[q], sw[q].col, x, y, c);
      else c = put_ball(sb[q - nWl], x, y, c);
    }
    out[(size_t)y * RL.W + x] = c;
  }
}

// ------------------------------------------------------------------------------------
// K3 fast path: walls inserted before balls (every DataSaver world).
//
// One warp owns one 64x64 tile (a whole frame on the 64x64 canvas) at a time and builds it in
// its private 16 KB of shared memory:
//   pass 1  background (white + walls): lane -> 4 consecutive pixels of a row (one 128-bit
//           st.shared, two rows per warp store, conflict-free).  The 4-pixel pattern only
//           changes where a wall's y range starts or ends, so it is rebuilt at those rows only;
//   pass 2  ball sprites from the atlas (L1-resident), OVER-composited in place, one ball at a
//           time in insertion order;
//   pass 3  the finished tile leaves through the TMA: one cp.async.bulk shared->global of the
//           whole frame when the tile is the frame (16 KB contiguous), else one bulk copy per
//           tile row.  The warp starts on its next tile's geometry while the copy engine reads
//           the buffer, and only waits (cp.async.bulk.wait_group.read) before overwriting it.
// HBM sees exactly one full-line write per frame byte and a few hundred bytes of state reads.
// ------------------------------------------------------------------------------------
// Warps per render CTA: as many as fit next to PPB_RENDER_SMEM_SPARE bytes (room for two advance CTAs beside the last
// renders of a chunk and for the shared-memory copy of a small sprite atlas; see render_warps()); the per-warp
// metadata is sized by the batch's actual wall / ball counts.
#define PPB_RENDER_MAX_WARPS 12
#define PPB_RENDER_SMEM_SPARE (2 * (11264 + 1024) + 1024)
struct RTBall {      // one ball inside the current tile, already clipped to it (32 bytes)
  uint32_t tex;      // atlas element index of the first visible sprite pixel
  uint32_t dst_sz;   // tile element offset of that pixel | sprite side << 16
  uint32_t aw_ah;    // visible width | visible height << 16   (0 = nothing to draw)
  float inv_aw;      // 1 / visible width
  uint32_t dti, ddi; // atlas / tile index step for +32 pixels (dq rows + dr columns)
  uint32_t wti_wdi;  // extra atlas | tile << 16 step when the column wraps into the next row
  uint32_t dq_dr;    // 32 / aw | 32 % aw << 16
};
// how pass 1 paints a wall inside the current tile
#define PPB_WM_SKIP 0    // does not touch the tile
#define PPB_WM_SOLID 1   // axis-aligned, integral edges, opaque colour: every pixel is either the colour or untouched
#define PPB_WM_FAST 2    // axis-aligned with integral y edges: coverage depends on the column only
#define PPB_WM_SLOW 3    // anything else: coverage evaluated per pixel
// per-warp shared memory: the tile, then metadata arrays sized by (n_walls, n_balls)
struct RenderWarpSmem {
  uint32_t *tile;      // [PPB_TILE * PPB_TILE]  16 KB, source of the bulk store
  int4 *wpaint;        // [n_walls] {PPB_WM_*, colour, first tile row | rows << 16, first column group | groups << 16}
  int4 *wspan;         // [n_walls] {tile-local pixel columns x0, x1 (SOLID), 1 / column groups as float bits, items}
  SWallGeom *w;        // [n_walls]
  RTBall *b;           // [n_balls]
};
__host__ __device__ inline size_t render_warp_meta_bytes(int n_walls, int n_balls) {
  size_t meta = (size_t)n_walls * (sizeof(int4) + sizeof(int4) + sizeof(SWallGeom)) + (size_t)n_balls * sizeof(RTBall);
  return (meta + 15) & ~(size_t)15;
}
// The CTA's shared memory: the tiles of all warps first (16 KB each, so every tile is 1024-byte aligned -- what the
// swizzled tensor-map store needs), then the metadata of all warps.
__host__ __device__ inline size_t render_warp_smem_bytes(int n_walls, int n_balls) {
  return (size_t)PPB_TILE * PPB_TILE * 4 + render_warp_meta_bytes(n_walls, n_balls);
}
// warps per CTA for a batch layout: fill the 227 KB, minus the spare for co-resident collide CTAs
inline int render_warps(int n_walls, int n_balls) {
  const size_t per = render_warp_smem_bytes(n_walls, n_balls);
  int w = (int)((227 * 1024 - PPB_RENDER_SMEM_SPARE) / per);
  if (w > PPB_RENDER_MAX_WARPS) w = PPB_RENDER_MAX_WARPS;
  static const int forced = getenv("PPB_RENDER_WARPS") ? atoi(getenv("PPB_RENDER_WARPS")) : 0;   // developer switch (A/B timing)
  if (forced > 0) {
    w = forced;
    const int fit = (int)((227 * 1024) / per);
    if (w > fit) w = fit;
  }
  return w < 1 ? 1 : w;
}

__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
// One 2-D tensor-map store of a 128-byte-swizzled tile: rows of 128 bytes (32 pixels), `row0` = first row of the box
// in the tensor (frames seen as an array of 128-byte lines).
__device__ __forceinline__ void bulk_tensor_store_2d(const CUtensorMap *tmap, int row0, const void *ssrc) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(0), "r"(row0),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc))
               : "memory");
}
// 128-byte swizzle of the copy engine (CU_TENSOR_MAP_SWIZZLE_128B): the 16-byte chunk index inside a 128-byte line is
// XORed with the line index modulo 8.  w = pixel (32-bit word) index in a 1024-byte aligned tile, ch = 16-byte chunk index.
template <bool SWZ> __device__ __forceinline__ uint32_t swz_w(uint32_t w) { return SWZ ? (w ^ ((w >> 3) & 0x1cu)) : w; }
template <bool SWZ> __device__ __forceinline__ int swz_ch(int ch) { return SWZ ? (ch ^ ((ch >> 3) & 7)) : ch; }
// ---- L2 policies: the render state is read evict-last, the frames are written evict-first (see load_world) ----
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_hint_v4(const uint4 *a, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_hint(const float2 *a, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ double2 ld_hint(const double2 *a, uint64_t pol) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_hint(const float *a, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void bulk_tensor_store_2d_hint(const CUtensorMap *tmap, int row0, const void *ssrc, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tmap), "r"(0),
               "r"(row0), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ int round_half_away_to_int(float v) { return (int)roundf(v); }
__device__ __forceinline__ int round_half_away_to_int(double v) { return (int)round(v); }

// src OVER dst without the white-canvas shortcut (branch-free; a transparent src leaves dst unchanged)
__device__ __forceinline__ uint32_t over_nb(uint32_t src, uint32_t dst) { return src + mul_un8x4(dst, 255u - (src >> 24)); }

// WallRC of every wall of worlds [world0, world0 + count): walls never move, so the bounding boxes,
// paste rectangles and row ranges the render kernel needs are derived once per ppb_set_walls / ppb_set_styles
template <typename T>
__global__ void wall_cache_kernel(StatePtrs S, int n_walls, int n_balls, RenderLayout RL, int world0, int count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)count * n_walls) return;
  const int q = (int)(i % n_walls);
  const long long idx = (long long)world0 * n_walls + i;
  const long long wld = world0 + i / n_walls;
  SWall s;
  setup_wall<T>(s, reinterpret_cast<const T *>(S.wall) + idx * 8, RL.wall[q]);
  WallRC rc;
  rc.x0 = s.x0; rc.x1 = s.x1; rc.y0 = s.y0; rc.y1 = s.y1;
  rc.bx0 = s.bx0; rc.bx1 = s.bx1; rc.by0 = s.by0; rc.by1 = s.by1;
  rc.lo = s.lo; rc.hi = s.hi;
  rc.flags = (s.axis ? 1 : 0) | (s.fast ? 2 : 0);
  rc.col = s.col;
  reinterpret_cast<WallRC *>(reinterpret_cast<char *>(S.wall_rc) + (size_t)wld * render_rec_stride(n_walls, n_balls))[q] = rc;
}

// BAND = false: 64 x 64 pixel tiles (the whole frame on the 64-pixel canvas).
// BAND = true : canvases of any other width W (W % 4 == 0, W <= 4096): a tile is a band of 4096 / W whole frame
//               rows, stored in shared memory with row stride W -- it is contiguous in the frame, so it leaves as
//               ONE linear bulk copy whatever the width.  (Square tiles of a wide canvas leave as 64 row pieces of
//               256 bytes; with W = 700 every other row starts 16 bytes off a 32-byte sector and the partial-sector
//               writes halve the store rate: 2.1 TB/s against 3.9 TB/s at W = 704.)
// MODE 2      : the 64-pixel-wide canvas of at most 64 rows (the whole frame is one tile).  The tile is kept in the copy
//               engine's 128-byte swizzle and leaves through a 2-D tensor map (frames = an array of 128-byte lines,
//               box = 32 pixels x 2 H lines): rows of the tile no longer alias on the same banks, so the 8-pixel-wide
//               sprites and the tall walls stop replaying their shared-memory accesses 4-8 times.
#define PPB_RM_TILE 0
#define PPB_RM_BAND 1
#define PPB_RM_SWZ 2
// SATL        : the sprite atlas (a few KB when the batch draws one or two radii) is copied into the CTA's shared memory
//               once per launch and the compositing loop reads its texels from there: one address instruction and a
//               shared-memory latency per texel instead of a 64-bit address chain and an L1 round trip.
template <typename T, int MODE, bool SATL = false>
__global__ void __maxnreg__(120) render_tile_kernel(StatePtrs S, Layout L, RenderLayout RL, const uint32_t *atlas,
                                                    uint8_t *frames, const __grid_constant__ CUtensorMap tmap,
                                                    uint32_t *tile_ctr, uint32_t kChunk) {
  typedef typename Vec2<T>::type T2;
  constexpr bool BAND = MODE == PPB_RM_BAND;
  constexpr bool SWZ = MODE == PPB_RM_SWZ;
  extern __shared__ __align__(1024) unsigned char render_smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;
  RenderWarpSmem sm;
  {
    unsigned char *base = render_smem_raw + (size_t)n_warps * PPB_TILE * PPB_TILE * 4 +
                          (size_t)warp * render_warp_meta_bytes(L.n_walls, L.n_balls);
    sm.tile = reinterpret_cast<uint32_t *>(render_smem_raw + (size_t)warp * PPB_TILE * PPB_TILE * 4);
    sm.wpaint = reinterpret_cast<int4 *>(base);
    sm.wspan = sm.wpaint + L.n_walls;                                      // 16-byte records first
    sm.b = reinterpret_cast<RTBall *>(sm.wspan + L.n_walls);
    sm.w = reinterpret_cast<SWallGeom *>(sm.b + L.n_balls);                // 4-byte aligned members only
  }
  if (SWZ && (__cvta_generic_to_shared(render_smem_raw) & 1023u) != 0) __trap();   // the swizzle is a function of the address
  // the atlas copy sits behind the tiles and the metadata of all warps
  const uint32_t *satlas = reinterpret_cast<const uint32_t *>(
      render_smem_raw + (size_t)n_warps * render_warp_smem_bytes(L.n_walls, L.n_balls));
  if (SATL) {
    uint32_t *dstw = reinterpret_cast<uint32_t *>(
        render_smem_raw + (size_t)n_warps * render_warp_smem_bytes(L.n_walls, L.n_balls));
    const int words = (L.n_balls * RL.sprite_slot_stride) >> 2;
    for (int i = threadIdx.x; i < words; i += blockDim.x) dstw[i] = __ldg(atlas + i);
    __syncthreads();
  }
  auto texel = [&](uint32_t idx) -> uint32_t { return SATL ? satlas[idx] : __ldg(atlas + idx); };
  // pixels per tile row (row stride in the buffer; a band pads it to a multiple of 4) and rows per tile
  const int RS = BAND ? ((RL.W + 3) & ~3) : PPB_TILE;
  const int RH = BAND ? (PPB_TILE * PPB_TILE) / RS : PPB_TILE;
  const int RS4 = RS >> 2;                                            // 4-pixel groups per tile row
  const int tiles_x = BAND ? 1 : (RL.W + PPB_TILE - 1) / PPB_TILE;
  const int tiles_y = (RL.H + RH - 1) / RH;
  const int tiles = tiles_x * tiles_y;
  const int nB = L.n_balls, nWl = L.n_walls;
  const long long total = (long long)S.n_worlds * tiles;
  const bool whole_rows = (BAND && RS == RL.W) || (RL.W == PPB_TILE);   // the tile is contiguous in the frame
  const bool bulk_ok = ((RL.W & 3) == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15u) == 0);
  // raw state of the world of the current / next tile, held in registers: lane q < nWl keeps wall q's
  // WallRC, lane b keeps ball b's position and radius.  The next tile's loads are
  // issued before the current tile is drawn, so their latency hides behind the drawing.
  uint4 rc0 = make_uint4(0, 0, 0, 0), rc1 = rc0;
  uint4 rc2 = rc0;   // every loaded word is consumed below: a dead destination register would be
                     // recycled by ptxas and stall its next writer on the load (128-bit scoreboard)
  T2 bp;
  T brad = T(-1);
  bp.x = bp.y = T(0);
  // small-sprite walk cached per lane: key = (visible width, height, sprite side), offsets of the lane's two pixels
  uint32_t sp_key = 0u, sp_t1 = 0xffffffffu, sp_t2 = 0xffffffffu, sp_d1 = 0u, sp_d2 = 0u;
  // Tiles are handed out dynamically: a warp's first tile comes from its grid index, every further one from a global
  // ticket (tile_ctr[0]); the SMs do not all drain their stores at the same rate (ncu: 595 k - 740 k active cycles per
  // scheduler under an equal static split), and the launch ends when the slowest one does.  The ticket for the tile
  // after next is drawn one iteration ahead, so its latency and the next tile's state loads hide behind the drawing.
  const long long job0 = (long long)blockIdx.x * n_warps + warp;
  const long long n_static = (long long)gridDim.x * n_warps;
  // (the atomic's result stays in lane 0's register until the next iteration broadcasts it: a shuffle right behind the
  // atomic would stall the warp for the round trip to L2)
  long long static_next = job0;   // (developer switch: tile_ctr == nullptr = the static round-robin split)
  // One atomic per tile on one address is more than an L2 slice takes (380 us per launch), and a ticket worth several
  // CONSECUTIVE frames is slow for another reason: frames that are neighbours in memory should be written at about the
  // same time, by different warps (tickets of 2 / 4 / 8 consecutive frames: 345 / 357 / 371 us).  So a ticket t of the
  // first phase is worth kChunk tiles that are A / kChunk apart -- the grid then fills kChunk windows of the frame array in
  // order, like the static split fills one -- and the last 3 tiles per warp are handed out one by one (no warp is left
  // working alone on a long ticket at the end).
  uint32_t ticket_raw = 0u, chunk_left = 0u;
  long long chunk_job = 0;
  const long long n_dyn = total > n_static ? total - n_static : 0;
  const long long A_per = kChunk > 1u && n_dyn > 3 * n_static ? (n_dyn - 3 * n_static) / kChunk : 0;   // tickets of phase A
  const long long A = A_per * kChunk;
  auto request_ticket = [&]() {
    if (!tile_ctr) { static_next += n_static; return; }
    if (chunk_left == 0u && lane == 0) ticket_raw = atomicAdd(tile_ctr, 1u);
  };
  // State prefetcher (one tile per world only).  The per-tile state reads are scattered single lines that interrupt HBM's
  // write stream (see load_world); the tickets move through kChunk windows of consecutive worlds, so the warp that draws
  // a ticket number divisible by kPrefetchTickets pulls the records and positions of the NEXT kPrefetchTickets worlds of
  // every window into L2 as a few long contiguous bursts.  About half of the later loads then hit L2 (the windows drawn
  // late in a ticket have been evicted again by the frame stream): 336 -> 327 us per 131,072 frames.
  constexpr long long kPrefetchTickets = 256;
  auto prefetch_bytes = [&](const char *b, size_t bytes) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(b) & ~(uintptr_t)127;
    const size_t lines = (reinterpret_cast<uintptr_t>(b) + bytes - a0 + 127) >> 7;
    for (size_t i = lane; i < lines; i += 32) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a0 + (i << 7)));
  };
  auto prefetch_worlds = [&](long long w0, long long n) {
    if (w0 + n > S.n_worlds) n = S.n_worlds - w0;
    if (n <= 0) return;
    const size_t rs = render_rec_stride(L.n_walls, L.n_balls);
    prefetch_bytes(reinterpret_cast<const char *>(S.wall_rc) + (size_t)w0 * rs, (size_t)n * rs);
    prefetch_bytes(reinterpret_cast<const char *>(S.pos) + (size_t)w0 * L.n_balls * sizeof(T2), (size_t)n * L.n_balls * sizeof(T2));
  };
  const bool prefetching = tiles == 1 && tile_ctr != nullptr && A_per > 0;
  if (prefetching && blockIdx.x == 0 && warp < (int)kChunk)
    prefetch_worlds(n_static + (long long)warp * A_per, kPrefetchTickets < A_per ? kPrefetchTickets : A_per);
  auto ticket_job = [&]() -> long long {
    if (!tile_ctr) return static_next;
    if (chunk_left == 0u) {
      const long long t = (long long)__shfl_sync(0xffffffffu, ticket_raw, 0);
      if (prefetching && t < A_per && (t & (kPrefetchTickets - 1)) == 0) {
        const long long t1 = t + kPrefetchTickets;
        if (t1 < A_per) {
          const long long n = (A_per - t1) < kPrefetchTickets ? (A_per - t1) : kPrefetchTickets;
          for (uint32_t k = 0; k < kChunk; ++k) prefetch_worlds(n_static + t1 + (long long)k * A_per, n);
        }
        if (t1 + kPrefetchTickets > A_per) prefetch_worlds(n_static + A, total - n_static - A);   // the single-tile tickets of the end
      }
      if (t < A_per) { chunk_job = n_static + t; chunk_left = kChunk; }
      else { chunk_job = n_static + A + (t - A_per); chunk_left = 1u; }
    }
    --chunk_left;
    const long long j = chunk_job;
    chunk_job += A_per;
    return j;
  };
  const size_t rec_stride = render_rec_stride(nWl, nB);
  // One contiguous record per world (walls, radii) and its ball positions.  L2-missing reads that trickle in during the
  // launch pull HBM out of its write stream (measured on B200, 131,072 frames, kernel without its paint loops: 298 us
  // without these loads, 327 us with a single 16-byte read per frame, 337 - 345 us with all of them, whatever the warp
  // count; two worlds per read, or one contiguous read per 12 worlds, changed nothing): they are read with an evict-last
  // L2 policy and the frames leave with an evict-first one, which is worth 8 us.
  const uint64_t pol_keep = l2_policy_evict_last();
  auto load_world = [&](int wn) {
    const char *rec = reinterpret_cast<const char *>(S.wall_rc) + (size_t)wn * rec_stride;
    if (lane < nWl) {
      const uint4 *src = reinterpret_cast<const uint4 *>(rec) + 3 * lane;
      rc0 = ld_hint_v4(src, pol_keep); rc1 = ld_hint_v4(src + 1, pol_keep); rc2 = ld_hint_v4(src + 2, pol_keep);
    }
    if (lane < nB) {
      bp = ld_hint(reinterpret_cast<const T2 *>(S.pos) + ((long long)wn * nB + lane), pol_keep);
      brad = (T)ld_hint(reinterpret_cast<const float *>(rec + (size_t)nWl * sizeof(WallRC)) + lane, pol_keep);
    }
  };
  if (job0 < total) load_world((tiles == 1) ? (int)job0 : (int)(job0 / tiles));
  if (job0 < total) request_ticket();
  for (long long job = job0; job < total;) {
    int w = (int)job, tx0 = 0, ty0 = 0;
    if (tiles != 1) {
      w = (int)(job / tiles);
      const int tile = (int)(job - (long long)w * tiles);
      const int tyi = tile / tiles_x;
      tx0 = (tile - tyi * tiles_x) * RS;
      ty0 = tyi * RH;
    }
    const int y_end = min(ty0 + RH, RL.H);
    const int x_end = min(tx0 + RS, RL.W);
    const int th = y_end - ty0, tw = x_end - tx0;
    // ---- geometry of this tile from the registers, then the loads of the next tile's world ----
    const uint4 c0 = rc0, c1 = rc1, c2 = rc2;
    const T2 cbp = bp;
    const T cbrad = brad;
    const long long nj = ticket_job();   // requested one iteration ago
    if (nj < total) {
      const int wn = (tiles == 1) ? (int)nj : (int)(nj / tiles);
      if (wn != w) load_world(wn);
      request_ticket();
    }
    if (lane < nWl) {
      const float wx0 = __uint_as_float(c0.x), wx1 = __uint_as_float(c0.y);
      const int bx0 = (int)c1.x, bx1 = (int)c1.y, lo = (int)c2.x, hi = (int)c2.y, flags = (int)c2.z;
      const uint32_t col = c2.w;
      // pixels of this tile the wall can touch
      const int px0 = max(max(bx0, (int)floorf(wx0)), tx0), px1 = min(min(bx1, (int)ceilf(wx1)), x_end);
      const int r0 = max(lo, ty0), r1 = min(hi, y_end);
      int mode = PPB_WM_SLOW;
      if (px0 >= px1 || r0 >= r1) mode = PPB_WM_SKIP;
      else if (flags & 2)
        mode = ((col >> 24) == 255u && floorf(wx0) == wx0 && floorf(wx1) == wx1) ? PPB_WM_SOLID : PPB_WM_FAST;
      const int g0 = (px0 - tx0) >> 2, g1 = (px1 - tx0 + 3) >> 2;
      sm.wpaint[lane] = make_int4(mode, (int)col, (r0 - ty0) | ((r1 - r0) << 16), g0 | ((g1 - g0) << 16));
      // the per-wall loop constants are derived here, by the wall's own lane, off the painting loop's critical path
      sm.wspan[lane] = make_int4(px0 - tx0, px1 - tx0, __float_as_int(__frcp_rn((float)max(g1 - g0, 1))), (r1 - r0) * (g1 - g0));
      if (mode >= PPB_WM_FAST) {
        SWallGeom g;
        g.x0 = wx0; g.x1 = wx1; g.y0 = __uint_as_float(c0.z); g.y1 = __uint_as_float(c0.w);
        g.bx0 = bx0; g.bx1 = bx1; g.by0 = (int)c1.z; g.by1 = (int)c1.w;
        g.axis = flags & 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) g.v[i] = 0.f;
        if (mode == PPB_WM_SLOW) {
          const T *wv = reinterpret_cast<const T *>(S.wall) + ((long long)w * nWl + lane) * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) g.v[i] = (float)wv[i];
        }
        sm.w[lane] = g;
      }
    }
    uint32_t my_dst_sz = 0u, my_aw_ah = 0u;
    if (lane < nB) {
      const int b = lane;
      RTBall rec;
      rec.tex = 0u; rec.dst_sz = 0u; rec.aw_ah = 0u; rec.inv_aw = 0.f;
      rec.dti = rec.ddi = rec.wti_wdi = rec.dq_dr = 0u;
      const int r = (int)cbrad;
      if (cbrad >= T(0) && r >= RL.r_min && r <= RL.r_max) {
        const int off = r + RL.s_thick[b];                      // floor(sz/2), primitives.py:420
        const int sz = 2 * off;
        const int sx0 = round_half_away_to_int(cbp.x) - off;    // x_asint (geometry.py:25-29)
        const int sy0 = round_half_away_to_int(cbp.y) - off;
        const int ax0 = max(sx0, tx0), ay0 = max(sy0, ty0);
        const int aw = min(sx0 + sz, x_end) - ax0, ah = min(sy0 + sz, y_end) - ay0;
        if (aw > 0 && ah > 0) {
          rec.tex = (uint32_t)(((size_t)b * RL.sprite_slot_stride + (size_t)(r - RL.r_min) * RL.sprite_stride) >> 2) +
                    (uint32_t)((ay0 - sy0) * sz + (ax0 - sx0));
          rec.dst_sz = (uint32_t)((ay0 - ty0) * RS + (ax0 - tx0)) | ((uint32_t)sz << 16);
          rec.aw_ah = (uint32_t)aw | ((uint32_t)ah << 16);
          const int dq = 32 / aw, dr = 32 - dq * aw;
          rec.inv_aw = 1.0f / (float)aw;
          rec.dti = (uint32_t)(dq * sz + dr);
          rec.ddi = (uint32_t)(dq * RS + dr);
          rec.wti_wdi = (uint32_t)(sz - aw) | ((uint32_t)(RS - aw) << 16);
          rec.dq_dr = (uint32_t)dq | ((uint32_t)dr << 16);
        }
      }
      sm.b[b] = rec;
      my_dst_sz = rec.dst_sz;
      my_aw_ah = rec.aw_ah;
    }
    // Which balls are composited together with their successor (pass 2)?  Decided here by every ball's own lane:
    // same visible geometry, small sprite, pixel boxes apart.  Bit b of pair_mask: ball b pairs with ball b + 1.
    unsigned pair_mask;
    {
      const uint32_t n_dst_sz = __shfl_down_sync(0xffffffffu, my_dst_sz, 1), n_aw_ah = __shfl_down_sync(0xffffffffu, my_aw_ah, 1);
      bool pair_ok = false;
      if (lane + 1 < nB && my_aw_ah != 0u && n_aw_ah == my_aw_ah && (n_dst_sz >> 16) == (my_dst_sz >> 16)) {
        const int aw = (int)(my_aw_ah & 0xffffu), ah = (int)(my_aw_ah >> 16);
        const int dst = (int)(my_dst_sz & 0xffffu), dstn = (int)(n_dst_sz & 0xffffu);
        const int dsty = BAND ? dst / RS : (dst >> 6), dstny = BAND ? dstn / RS : (dstn >> 6);
        const int ddx = abs((dst - dsty * RS) - (dstn - dstny * RS)), ddy = abs(dsty - dstny);
        pair_ok = aw * ah <= 64 && (ddx >= aw || ddy >= ah);
      }
      pair_mask = __ballot_sync(0xffffffffu, pair_ok);
    }
    bulk_wait_read0();   // the copy engine is done reading the tile buffer
    __syncwarp();
    // ---- pass 1: background.  White everywhere (rows / columns of the 64x64 buffer outside the
    //      canvas are written too, they never leave shared memory), then every wall is painted over
    //      the pixels of its bounding box in insertion order: lane -> (row, 4-pixel group) items ----
    {
      uint4 *t4 = reinterpret_cast<uint4 *>(sm.tile) + lane;
      const uint4 white = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
#pragma unroll 8
      for (int i = 0; i < PPB_TILE / 2; ++i) t4[i * 32] = white;
    }
    __syncwarp();
    for (int q = 0; q < nWl; ++q) {
      const int4 h = sm.wpaint[q];
      if (h.x == PPB_WM_SKIP) continue;
      const uint32_t col = (uint32_t)h.y;
      const int r0 = h.z & 0xffff, g0 = h.w & 0xffff, ncg = h.w >> 16;
      const int4 span = sm.wspan[q];
      const int items = span.w;
      const float inv_ncg = __int_as_float(span.z);
      uint4 *tile4 = reinterpret_cast<uint4 *>(sm.tile);
      const int ch0 = r0 * RS4 + g0;
      if (h.x == PPB_WM_SOLID) {
        const uint32_t width = (uint32_t)(span.y - span.x);
        // two items per trip (i and i + 32: different groups of the tile, so the two read-modify-write chains are
        // independent and their shared-memory latencies overlap -- this loop is latency-bound, one chain per warp)
        for (int i = lane; i < items; i += 64) {
          const int i2 = i + 32;
          const bool has2 = i2 < items;
          const int r = (int)(((float)i + 0.5f) * inv_ncg);
          const int r2 = (int)(((float)i2 + 0.5f) * inv_ncg);
          const int c = i - r * ncg, c2 = i2 - r2 * ncg;
          uint4 *p = tile4 + swz_ch<SWZ>(ch0 + r * RS4 + c);
          uint4 *p2 = tile4 + swz_ch<SWZ>(ch0 + (has2 ? r2 * RS4 + c2 : 0));
          const uint32_t d = (uint32_t)(((g0 + c) << 2) - span.x);   // pixel k of the group is inside iff d + k < width
          const uint32_t d2 = (uint32_t)(((g0 + c2) << 2) - span.x);
          uint4 v = *p;
          uint4 v2 = has2 ? *p2 : make_uint4(0u, 0u, 0u, 0u);
          v.x = (d < width) ? col : v.x;
          v.y = (d + 1u < width) ? col : v.y;
          v.z = (d + 2u < width) ? col : v.z;
          v.w = (d + 3u < width) ? col : v.w;
          v2.x = (d2 < width) ? col : v2.x;
          v2.y = (d2 + 1u < width) ? col : v2.y;
          v2.z = (d2 + 2u < width) ? col : v2.z;
          v2.w = (d2 + 3u < width) ? col : v2.w;
          *p = v;
          if (has2) *p2 = v2;
        }
      } else {
        const SWallGeom &g = sm.w[q];
        for (int i = lane; i < items; i += 32) {
          const int r = (int)(((float)i + 0.5f) * inv_ncg);
          const int c = i - r * ncg;
          uint4 *p = tile4 + swz_ch<SWZ>(ch0 + r * RS4 + c);
          const int x = tx0 + ((g0 + c) << 2), y = ty0 + r0 + r;
          uint4 v = *p;
          if (h.x == PPB_WM_FAST) {
            // full row coverage: the mask of a pixel is its x-overlap with the wall
            uint32_t m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int xp = x + k;
              const float ox = fminf((float)(xp + 1), g.x1) - fmaxf((float)xp, g.x0);
              m[k] = (xp >= g.bx0 && xp < g.bx1 && ox > 0.f) ? (uint32_t)floorf(255.f * ox + 0.5f) : 0u;
            }
            v.x = over_nb(mul_un8x4(col, m[0]), v.x);
            v.y = over_nb(mul_un8x4(col, m[1]), v.y);
            v.z = over_nb(mul_un8x4(col, m[2]), v.z);
            v.w = over_nb(mul_un8x4(col, m[3]), v.w);
          } else {
            put_wall4(g, col, x, y, v.x, v.y, v.z, v.w);
          }
          *p = v;
        }
      }
      __syncwarp();
    }
    // ---- pass 2: ball sprites, composited in place in insertion order.  All lanes work on one
    //      sprite: lane -> pixel (sy, sx) of its visible part, then +32 pixels per iteration.
    //      Small sprites (<= 64 visible pixels: two per lane) take a short path whose per-lane pixel
    //      offsets only depend on (visible width, height, sprite side) -- the same for every unclipped
    //      ball of one radius -- and are kept in registers from ball to ball and tile to tile ----
    for (int b = 0; b < nB; ++b) {
      const uint4 ra = reinterpret_cast<const uint4 *>(&sm.b[b])[0];   // tex, dst_sz, aw_ah, inv_aw
      if (ra.z == 0u) continue;
      const int aw = (int)(ra.z & 0xffffu), ah = (int)(ra.z >> 16);
      const int sz = (int)(ra.y >> 16);
      if (aw * ah <= 64) {
        const uint32_t key = ra.z ^ (ra.y & 0xffff0000u) ^ 0x80000000u;   // (aw, ah, sz); never 0
        if (key != sp_key) {
          sp_key = key;
          const float inv = __uint_as_float(ra.w);
          const int y1 = (int)(((float)lane + 0.5f) * inv), x1 = lane - y1 * aw;
          const int y2 = (int)(((float)(lane + 32) + 0.5f) * inv), x2 = lane + 32 - y2 * aw;
          sp_t1 = (y1 < ah) ? (uint32_t)(y1 * sz + x1) : 0xffffffffu;
          sp_d1 = (uint32_t)(y1 * RS + x1);
          sp_t2 = (y2 < ah) ? (uint32_t)(y2 * sz + x2) : 0xffffffffu;
          sp_d2 = (uint32_t)(y2 * RS + x2);
        }
        const uint32_t dst = ra.y & 0xffffu;
        const bool v1 = sp_t1 != 0xffffffffu, v2 = sp_t2 != 0xffffffffu;
        const uint32_t a1 = swz_w<SWZ>(dst + sp_d1), a2 = swz_w<SWZ>(dst + sp_d2);
        // the next ball, when it has the same visible geometry and does not touch this one's pixels, is
        // composited in the same trip: four independent load -> blend -> store chains instead of two
        // (measured: step 0.459 -> 0.421 ms; grouping up to four sprites was slower again, 0.459 ms)
        {
          if ((pair_mask >> b) & 1u) {
            const uint2 rn = reinterpret_cast<const uint2 *>(&sm.b[b + 1])[0];   // tex, dst_sz of the partner
            const uint32_t dstn = rn.y & 0xffffu;
            const uint32_t n1 = swz_w<SWZ>(dstn + sp_d1), n2 = swz_w<SWZ>(dstn + sp_d2);
            const uint32_t t1 = v1 ? texel(ra.x + sp_t1) : 0u;
            const uint32_t t2 = v2 ? texel(ra.x + sp_t2) : 0u;
            const uint32_t u1 = v1 ? texel(rn.x + sp_t1) : 0u;
            const uint32_t u2 = v2 ? texel(rn.x + sp_t2) : 0u;
            const uint32_t d1 = v1 ? sm.tile[a1] : 0u;
            const uint32_t d2 = v2 ? sm.tile[a2] : 0u;
            const uint32_t e1 = v1 ? sm.tile[n1] : 0u;
            const uint32_t e2 = v2 ? sm.tile[n2] : 0u;
            if (v1) { sm.tile[a1] = over_nb(t1, d1); sm.tile[n1] = over_nb(u1, e1); }
            if (v2) { sm.tile[a2] = over_nb(t2, d2); sm.tile[n2] = over_nb(u2, e2); }
            __syncwarp();
            ++b;
            continue;
          }
        }
        const uint32_t t1 = v1 ? texel(ra.x + sp_t1) : 0u;
        const uint32_t t2 = v2 ? texel(ra.x + sp_t2) : 0u;
        const uint32_t d1 = v1 ? sm.tile[a1] : 0u;
        const uint32_t d2 = v2 ? sm.tile[a2] : 0u;
        if (v1) sm.tile[a1] = over_nb(t1, d1);
        if (v2) sm.tile[a2] = over_nb(t2, d2);
        __syncwarp();
        continue;
      }
      const uint4 rb = reinterpret_cast<const uint4 *>(&sm.b[b])[1];   // dti, ddi, wti_wdi, dq_dr
      int sy = (int)(((float)lane + 0.5f) * __uint_as_float(ra.w));
      int sx = lane - sy * aw;
      const int dq = (int)(rb.w & 0xffffu), dr = (int)(rb.w >> 16);
      uint32_t ti = ra.x + (uint32_t)(sy * sz + sx);                      // atlas element index
      uint32_t di = (ra.y & 0xffffu) + (uint32_t)(sy * RS + sx);          // tile element index
      const uint32_t wti = rb.z & 0xffffu, wdi = rb.z >> 16;
      // two pixels per trip (this lane's pixel and the one 32 further on): two independent
      // load -> blend -> store chains in flight
      while (sy < ah) {
        int sx2 = sx + dr, sy2 = sy + dq;
        uint32_t ti2 = ti + rb.x, di2 = di + rb.y;
        if (sx2 >= aw) { sx2 -= aw; ++sy2; ti2 += wti; di2 += wdi; }
        const bool has2 = sy2 < ah;
        const uint32_t t = texel(ti);
        const uint32_t t2 = has2 ? texel(ti2) : 0u;
        const uint32_t ai = swz_w<SWZ>(di), ai2 = swz_w<SWZ>(di2);
        const uint32_t d = sm.tile[ai];
        const uint32_t d2 = has2 ? sm.tile[ai2] : 0u;
        sm.tile[ai] = over_nb(t, d);
        if (has2) sm.tile[ai2] = over_nb(t2, d2);
        sx = sx2 + dr; sy = sy2 + dq; ti = ti2 + rb.x; di = di2 + rb.y;
        if (sx >= aw) { sx -= aw; ++sy; ti += wti; di += wdi; }
      }
      __syncwarp();
    }
    // ---- pass 3: tile -> HBM through the copy engine ----
    fence_async_smem();   // this lane's generic-proxy writes become visible to the async proxy
    __syncwarp();
    uint32_t *out = reinterpret_cast<uint32_t *>(frames) + (size_t)w * RL.W * RL.H;
    if (SWZ) {
      // (the launcher only picks this mode for 16-byte aligned frames of 64 x H pixels, H <= 64: one tile per frame)
      if (lane == 0) bulk_tensor_store_2d_hint(&tmap, w * (2 * RL.H), sm.tile, l2_policy_evict_first());
    } else if (BAND && !bulk_ok) {
      // Rows of this canvas are not 16-byte multiples (the reference's default arenaSz = 667) or the frame buffer
      // is not 16-byte aligned: the copy engine cannot take the band.  It is still one contiguous run of
      // th * W pixels in the frame, so the warp streams it out with 16-byte stores from the first aligned pixel
      // on (the four pixels of a store may sit on two rows of the padded buffer), scalar stores at both ends.
      const int Wp = RL.W, npx = th * Wp;
      uint32_t *gp = out + (size_t)ty0 * Wp;
      const int lead = min(npx, (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(gp) & 15u)) & 15u) >> 2));
      const int nvec = (npx - lead) >> 2;
      const float inv_w = __frcp_rn((float)Wp);
      if (lane < lead) {
        const int r = lane / Wp;
        gp[lane] = sm.tile[r * RS + (lane - r * Wp)];
      }
      for (int v = lane; v < nvec; v += 32) {
        const int p = lead + (v << 2);
        int r = (int)(((float)p + 0.5f) * inv_w);
        int c = p - r * Wp;
        if (c < 0) { --r; c += Wp; } else if (c >= Wp) { ++r; c -= Wp; }
        uint32_t q[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          q[k] = sm.tile[r * RS + c];
          if (++c == Wp) { c = 0; ++r; }
        }
        *reinterpret_cast<uint4 *>(gp + p) = make_uint4(q[0], q[1], q[2], q[3]);
      }
      {
        const int p = lead + (nvec << 2) + lane;
        if (p < npx) {
          const int r = p / Wp;
          gp[p] = sm.tile[r * RS + (p - r * Wp)];
        }
      }
    } else if (!bulk_ok) {
      // square tiles of a canvas whose rows the copy engine cannot take: row by row, coalesced 4-byte stores
      for (int r = 0; r < th; ++r) {
        uint32_t *orow = out + (size_t)(ty0 + r) * RL.W + tx0;
        for (int c = lane; c < tw; c += 32) orow[c] = sm.tile[r * RS + c];
      }
    } else if (whole_rows) {
      if (lane == 0) bulk_store_s2g(out + (size_t)ty0 * RL.W, sm.tile, (uint32_t)(th * RS) * 4u);
    } else {
      const uint32_t row_bytes = (uint32_t)tw * 4u;
      for (int r = lane; r < th; r += 32)
        bulk_store_s2g(out + (size_t)(ty0 + r) * RL.W + tx0, sm.tile + r * RS, row_bytes);
    }
    bulk_commit();
    job = nj;
  }
  // the last warp of the grid to get here re-arms the ticket counter for the next launch that uses it
  if (lane == 0 && tile_ctr) {
    __threadfence();
    if (atomicAdd(tile_ctr + 1, 1u) == (uint32_t)(n_static - 1)) {
      tile_ctr[0] = 0u;
      tile_ctr[1] = 0u;
      __threadfence();
    }
  }
  bulk_wait0();   // smem must stay valid until the copy engine has read it; writes complete before exit
}

}  // namespace ppb
