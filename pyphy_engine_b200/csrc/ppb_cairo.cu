// ppb_cairo.cu -- ball sprites the way the reference's Cairo calls produce them (get_ball_im,
// primitives.py:33-61:  arc + set_line_width + stroke, arc + fill  on an ARGB32 surface).
//
// The reference reaches libcairo through pycairo; neither is part of its tree.  This unit implements the
// published pipeline of cairo's image backend (1.12 - 1.16) for exactly those calls:
//   host   (once per ppb_set_styles, a few hundred vertices): cairo_arc's Bezier segments, 24.8 fixed point,
//          de Casteljau flattening at tolerance 0.1, the stroker's two offset contours, and the conversion of
//          every polygon edge to the scan converter's grid (15 sub-rows per pixel, x in 1/256 px);
//   device (one thread per sprite pixel): the coverage the "tor" scan converter accumulates for that pixel --
//          rows in which no edge starts, ends or crosses another are covered analytically per edge
//          (cell_list_render_edge), all other rows are sampled on their 15 sub-rows -- then
//          alpha = (17 * area + 256) >> 9 and the span compositing (lerp for opaque colours, OVER through the
//          mask otherwise).
// The tests' CPU checker restates the same pipeline as cairo structures it (edge walking into per-row cell
// lists); here every pixel evaluates its own closed form.
#include <cmath>
#include <cstdint>
#include <vector>

#include "ppb_render.cuh"
#include "ppb_launch.h"

namespace ppb {
namespace {

constexpr int kGridY = 15;                   // sub-rows per pixel row
constexpr int kFullArea = 2 * 256 * kGridY;  // coverage of a fully covered pixel
constexpr double kTolerance = 0.1;           // cairo's default path tolerance
constexpr int kMaxActive = 32;               // edges of one sprite polygon that can meet one pixel row

struct Fx2 { int32_t x, y; };                // a point in 24.8 fixed point
struct GridEdge {                            // polygon edge on the scan converter's grid (y in 1/15 px, x in 1/256 px)
  int32_t x1, y1, dx, dy;                    // x(y) = x1 + floor((y - y1) * dx / dy)
  int32_t ytop, ybot, dir, pad;              // active for ytop <= y < ybot (clamped to the surface); winding +-1
};

inline int32_t to_fixed(double v) { return (int32_t)std::nearbyint(v * 256.0); }   // ties to even, like cairo's magic add
inline double from_fixed(int32_t v) { return (double)v / 256.0; }

// ---- path construction (host) ------------------------------------------------------------------------------------
struct Cubic { Fx2 p[4]; };

double arc_max_angle(double tolerance_over_radius) {
  // largest angle a single Bezier may span for a given normalised error (cairo-arc.c's table, then the formula)
  static const double err[11] = {0.0185185185185185036127, 0.000272567143730179811158, 2.38647043651461047433e-05,
                                 4.2455377443222443279e-06, 1.11281001494389081528e-06, 3.72662000942734705475e-07,
                                 1.47783685574284411325e-07, 6.63240432022601149057e-08, 3.2715520137536980553e-08,
                                 1.73863223499021216974e-08, 9.81410988043554039085e-09};
  for (int i = 0; i < 11; ++i)
    if (err[i] < tolerance_over_radius) return M_PI / (i + 1);
  int i = 12;
  double angle, e;
  do {
    angle = M_PI / i++;
    e = 2.0 / 27.0 * std::pow(std::sin(angle / 4), 6) / std::pow(std::cos(angle / 4), 2);
  } while (e > tolerance_over_radius);
  return angle;
}

// arc(xc, yc, r, 0, 2 pi): two half turns, each cut into equal Beziers with h = 4/3 tan(angle / 4)
std::vector<Cubic> circle_path(double xc, double yc, double r) {
  std::vector<Cubic> out;
  Fx2 cur = {to_fixed(xc + r * std::cos(0.0)), to_fixed(yc + r * std::sin(0.0))};
  const double full = 2 * M_PI, mid = 0.0 + (full - 0.0) / 2.0;
  const double piece[2][2] = {{0.0, mid}, {mid, full}};
  for (int h = 0; h < 2; ++h) {
    double a = piece[h][0];
    const double b = piece[h][1];
    int n = (int)std::ceil(std::fabs(b - a) / arc_max_angle(kTolerance / r));
    const double step = (b - a) / n;
    for (int i = 0; i < n; ++i) {
      const double A = a, B = (i == n - 1) ? b : a + step;
      const double rsA = r * std::sin(A), rcA = r * std::cos(A), rsB = r * std::sin(B), rcB = r * std::cos(B);
      const double hh = 4.0 / 3.0 * std::tan((B - A) / 4.0);
      Cubic c;
      c.p[0] = cur;
      c.p[1] = {to_fixed(xc + rcA - hh * rsA), to_fixed(yc + rsA + hh * rcA)};
      c.p[2] = {to_fixed(xc + rcB + hh * rsB), to_fixed(yc + rsB - hh * rcB)};
      c.p[3] = {to_fixed(xc + rcB), to_fixed(yc + rsB)};
      out.push_back(c);
      cur = c.p[3];
      a += step;
    }
  }
  return out;
}

struct FlatPoint { Fx2 p; int32_t tx, ty; };   // a flattened vertex and the tangent cairo reports with it

inline Fx2 half(Fx2 a, Fx2 b) { return {a.x + ((b.x - a.x) >> 1), a.y + ((b.y - a.y) >> 1)}; }

double flatness2(const Cubic &c) {   // squared distance of the farther inner control point from the chord (cairo-spline.c)
  double bx = from_fixed(c.p[1].x - c.p[0].x), by = from_fixed(c.p[1].y - c.p[0].y);
  double cx = from_fixed(c.p[2].x - c.p[0].x), cy = from_fixed(c.p[2].y - c.p[0].y);
  if (c.p[0].x != c.p[3].x || c.p[0].y != c.p[3].y) {
    const double dx = from_fixed(c.p[3].x - c.p[0].x), dy = from_fixed(c.p[3].y - c.p[0].y);
    const double v = dx * dx + dy * dy;
    double u = bx * dx + by * dy;
    if (u <= 0) { } else if (u >= v) { bx -= dx; by -= dy; } else { bx -= u / v * dx; by -= u / v * dy; }
    u = cx * dx + cy * dy;
    if (u <= 0) { } else if (u >= v) { cx -= dx; cy -= dy; } else { cx -= u / v * dx; cy -= u / v * dy; }
  }
  const double be = bx * bx + by * by, ce = cx * cx + cy * cy;
  return be > ce ? be : ce;
}

// vertices after the first one, in path order: the start of every leaf of the halving tree (duplicates dropped),
// then the end point with the curve's final slope
void flatten(const Cubic &curve, std::vector<FlatPoint> &out) {
  std::vector<Cubic> stack{curve};
  Fx2 last = curve.p[0];
  while (!stack.empty()) {
    Cubic c = stack.back();
    stack.pop_back();
    if (flatness2(c) < kTolerance * kTolerance) {
      if (c.p[0].x != last.x || c.p[0].y != last.y) {
        out.push_back({c.p[0], c.p[1].x - c.p[0].x, c.p[1].y - c.p[0].y});
        last = c.p[0];
      }
      continue;
    }
    const Fx2 ab = half(c.p[0], c.p[1]), bc = half(c.p[1], c.p[2]), cd = half(c.p[2], c.p[3]);
    const Fx2 abbc = half(ab, bc), bccd = half(bc, cd), m = half(abbc, bccd);
    stack.push_back(Cubic{{m, bccd, cd, c.p[3]}});      // second half, drawn after ...
    stack.push_back(Cubic{{c.p[0], ab, abbc, m}});      // ... the first half
  }
  Fx2 from = curve.p[2];
  if (from.x == curve.p[3].x && from.y == curve.p[3].y)
    from = (curve.p[1].x == curve.p[3].x && curve.p[1].y == curve.p[3].y) ? curve.p[0] : curve.p[1];
  out.push_back({curve.p[3], curve.p[3].x - from.x, curve.p[3].y - from.y});
}

struct EdgeList {
  std::vector<GridEdge> e;
  int rows;   // surface height in pixels
  void add(Fx2 a, Fx2 b) {
    if (a.y == b.y) return;
    int dir = 1;
    if (a.y > b.y) { std::swap(a, b); dir = -1; }
    auto gy = [](int32_t v) { return (int32_t)(((int64_t)kGridY * v + 128) >> 8); };
    const int32_t top = gy(a.y), bot = gy(b.y), ymax = rows * kGridY;
    if (top >= bot || top >= ymax || bot <= 0) return;
    GridEdge g;
    g.x1 = a.x; g.y1 = top; g.dx = b.x - a.x; g.dy = bot - top;
    g.ytop = top > 0 ? top : 0;
    g.ybot = bot < ymax ? bot : ymax;
    g.dir = dir; g.pad = 0;
    e.push_back(g);
  }
  void loop(const std::vector<Fx2> &v) {
    for (size_t i = 0; i < v.size(); ++i) add(v[i], v[(i + 1) % v.size()]);
  }
};

struct Face { Fx2 left, at, right; double ux, uy; };
Face face_at(Fx2 p, int32_t tx, int32_t ty, double half_width) {
  double ux = from_fixed(tx), uy = from_fixed(ty);
  if (ux == 0.0 && uy == 0.0) { }
  else if (ux == 0.0) uy = uy > 0 ? 1.0 : -1.0;
  else if (uy == 0.0) ux = ux > 0 ? 1.0 : -1.0;
  else { const double m = std::sqrt(ux * ux + uy * uy); ux /= m; uy /= m; }
  const int32_t ox = to_fixed(-uy * half_width), oy = to_fixed(ux * half_width);
  return Face{{p.x + ox, p.y + oy}, p, {p.x - ox, p.y - oy}, ux, uy};
}

// fill polygon of the circle
void circle_fill_edges(const std::vector<Cubic> &path, EdgeList &E) {
  std::vector<FlatPoint> fp;
  for (const Cubic &c : path) flatten(c, fp);
  std::vector<Fx2> v{path[0].p[0]};
  for (const FlatPoint &q : fp) v.push_back(q.p);
  E.loop(v);
}
// stroke polygon of the (open) circle path: the stroker puts one face per flattened vertex on each side; where two
// Beziers meet and their faces differ by a fixed-point step, the outer side takes the new face point and the inner
// side passes through the path point; the butt caps close the two contours into one loop
void circle_stroke_edges(const std::vector<Cubic> &path, double line_width, EdgeList &E) {
  const double hw = line_width / 2.0;
  std::vector<Fx2> L, R;
  Face cur{};
  for (size_t i = 0; i < path.size(); ++i) {
    const Cubic &c = path[i];
    const Face f = face_at(c.p[0], c.p[1].x - c.p[0].x, c.p[1].y - c.p[0].y, hw);
    if (i == 0) { L.push_back(f.left); R.push_back(f.right); }
    else if (f.left.x != cur.left.x || f.left.y != cur.left.y || f.right.x != cur.right.x || f.right.y != cur.right.y) {
      if (cur.ux * f.uy - cur.uy * f.ux > 0) { R.push_back(f.right); L.push_back(cur.at); L.push_back(f.left); }
      else { L.push_back(f.left); R.push_back(cur.at); R.push_back(f.right); }
    }
    cur = f;
    std::vector<FlatPoint> fp;
    flatten(c, fp);
    for (const FlatPoint &q : fp) {
      if ((q.tx | q.ty) == 0) continue;
      cur = face_at(q.p, q.tx, q.ty, hw);
      L.push_back(cur.left);
      R.push_back(cur.right);
    }
  }
  std::vector<Fx2> ring(L);
  ring.insert(ring.end(), R.rbegin(), R.rend());
  E.loop(ring);
}

// polygons of every (slot, radius): span[(slot * nr + ri) * 4 + ...] = first stroke edge, stroke edges, first fill
// edge, fill edges.  Slots that share a stroke width share the geometry, but the lists are tiny.
void sprite_polygons(const RenderLayout &RL, int n_balls, std::vector<GridEdge> &edges, std::vector<int32_t> &span) {
  const int nr = RL.r_max - RL.r_min + 1;
  span.assign((size_t)n_balls * nr * 4, 0);
  for (int b = 0; b < n_balls; ++b)
    for (int ri = 0; ri < nr; ++ri) {
      const int r = RL.r_min + ri, sT = RL.s_thick[b];
      int32_t *sp = &span[((size_t)b * nr + ri) * 4];
      if (r <= 0) continue;
      const double c = (double)(r + sT);
      const std::vector<Cubic> path = circle_path(c, c, (double)r);
      EdgeList E;
      E.rows = 2 * (r + sT);
      if (sT > 0) circle_stroke_edges(path, (double)sT, E);
      sp[0] = (int32_t)edges.size(); sp[1] = (int32_t)E.e.size();
      edges.insert(edges.end(), E.e.begin(), E.e.end());
      E.e.clear();
      circle_fill_edges(path, E);
      sp[2] = (int32_t)edges.size(); sp[3] = (int32_t)E.e.size();
      edges.insert(edges.end(), E.e.begin(), E.e.end());
    }
}

// ---- scan conversion (device, per pixel) -----------------------------------------------------------------------------
__host__ __device__ __forceinline__ long long floor_div(long long a, long long b) {
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
__host__ __device__ __forceinline__ int edge_x(const GridEdge &e, int y) {
  return e.dx == 0 ? e.x1 : e.x1 + (int)floor_div((long long)(y - e.y1) * e.dx, e.dy);
}
// what one edge of a whole-row step adds to the coverage of pixel column c: xt / xb are its x at the top / bottom
// of the pixel row; sign +1 opens a covered interval, -1 closes it
__host__ __device__ inline int full_row_edge_area(int xt, int xb, int sign, int c) {
  int ix1 = xt >> 8, ix2 = xb >> 8, fx1 = xt & 255, fx2 = xb & 255;
  if (ix1 == ix2) {
    if (c < ix1) return 0;
    return 512 * sign * kGridY - (c == ix1 ? sign * (fx1 + fx2) * kGridY : 0);
  }
  int dx = xb - xt, dy = kGridY;
  if (dx < 0) {
    int t = ix1; ix1 = ix2; ix2 = t;
    t = fx1; fx1 = fx2; fx2 = t;
    dx = -dx; sign = -sign; dy = -kGridY;
  }
  if (c < ix1) return 0;
  if (c > ix2) return 512 * sign * dy;
  // sub-rows (whole ones) the edge has crossed by the right border of column ix1 + k
  auto crossed = [&](int k) { return (int)floor_div((long long)((256 - fx1) + 256 * k) * dy, dx); };
  if (c == ix1) { const int y = crossed(0); return 512 * sign * y - sign * y * (256 + fx1); }
  if (c < ix2) { const int y = crossed(c - ix1), yp = crossed(c - ix1 - 1); return 512 * sign * y - sign * (y - yp) * 256; }
  return 512 * sign * dy - sign * (dy - crossed(ix2 - ix1 - 1)) * fx2;
}

// coverage (0 .. kFullArea) of the polygon E[0..n) under the non-zero rule on pixel (px, py); -1 = too many edges
__host__ __device__ inline int pixel_area(const GridEdge *E, int n, int px, int py) {
  const int y0 = py * kGridY;
  int idx[kMaxActive];
  int m = 0;
  bool starts_here = false, whole = true;
  for (int i = 0; i < n; ++i) {
    const int yt = E[i].ytop, yb = E[i].ybot;
    if (yb <= y0 || yt >= y0 + kGridY) continue;
    if (yt >= y0) starts_here = true;
    if (yb < y0 + kGridY) whole = false;
    if (m == kMaxActive) return -1;
    idx[m++] = i;
  }
  if (m == 0) return 0;
  whole = whole && !starts_here;
  int xs[kMaxActive], xe[kMaxActive], dr[kMaxActive];
  int area = 0;
  if (whole) {
    for (int k = 0; k < m; ++k) {
      const GridEdge &e = E[idx[k]];
      const int xt = edge_x(e, y0), xb = edge_x(e, y0 + kGridY), d = e.dir;
      int j = k;
      while (j > 0 && (xs[j - 1] > xt || (xs[j - 1] == xt && xe[j - 1] > xb))) { xs[j] = xs[j - 1]; xe[j] = xe[j - 1]; dr[j] = dr[j - 1]; --j; }
      xs[j] = xt; xe[j] = xb; dr[j] = d;
    }
    for (int k = 1; k < m; ++k)
      if (xe[k] <= xe[k - 1]) { whole = false; break; }        // two edges meet or cross inside this row
  }
  if (whole) {
    int k = 0;
    while (k < m) {
      const int left = k;
      int winding = dr[k];
      while (true) {
        ++k;
        if (k >= m) break;
        winding += dr[k];
        if (winding == 0 && (k + 1 >= m || xs[k + 1] != xs[k])) break;
      }
      if (k >= m) break;
      area += full_row_edge_area(xs[left], xe[left], +1, px) + full_row_edge_area(xs[k], xe[k], -1, px);
      ++k;
    }
  } else {
    const int pl = px * 256, pr = pl + 256;
    for (int sub = 0; sub < kGridY; ++sub) {
      const int y = y0 + sub;
      int na = 0;
      for (int k = 0; k < m; ++k) {
        const GridEdge &e = E[idx[k]];
        if (e.ytop > y || y >= e.ybot) continue;
        const int x = edge_x(e, y), d = e.dir;
        int j = na++;
        while (j > 0 && xs[j - 1] > x) { xs[j] = xs[j - 1]; dr[j] = dr[j - 1]; --j; }
        xs[j] = x; dr[j] = d;
      }
      int winding = 0, open = 0;
      for (int k = 0; k < na; ++k) {
        if (winding == 0) open = xs[k];
        winding += dr[k];
        if (winding == 0) {
          const int a = open > pl ? open : pl, b = xs[k] < pr ? xs[k] : pr;
          if (b > a) area += 2 * (b - a);
        }
      }
    }
  }
  return area < 0 ? 0 : (area > kFullArea ? kFullArea : area);
}
__host__ __device__ __forceinline__ uint32_t area_to_alpha(int area) { return (uint32_t)((area + (area << 4) + 256) >> 9); }

// one span pixel of the image compositor: opaque colours are blended in place (multiplies biased by 0x7f),
// translucent ones go through pixman's OVER with the coverage as mask
__host__ __device__ __forceinline__ uint32_t mul8_7f(uint32_t a, uint32_t b) { const uint32_t t = a * b + 0x7fu; return ((t >> 8) + t) >> 8; }
__host__ __device__ __forceinline__ uint32_t mul8_80(uint32_t a, uint32_t b) { const uint32_t t = a * b + 0x80u; return ((t >> 8) + t) >> 8; }
__host__ __device__ inline uint32_t paint_through_mask(uint32_t col, uint32_t a, uint32_t dst) {
  if (a == 0u) return dst;
  if ((col >> 24) != 255u) {                         // (colour IN mask) OVER dst, pixman's 0x80-biased multiplies
    const uint32_t ia = 255u - mul8_80(col >> 24, a);
    uint32_t out = 0u;
    for (int c = 0; c < 4; ++c) {
      uint32_t v = mul8_80((col >> (8 * c)) & 0xffu, a) + mul8_80((dst >> (8 * c)) & 0xffu, ia);
      v = v > 255u ? 255u : v;
      out |= v << (8 * c);
    }
    return out;
  }
  if (a == 255u) return col;
  uint32_t out = 0u;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t s = (col >> (8 * c)) & 0xffu, d = (dst >> (8 * c)) & 0xffu;
    uint32_t v = mul8_7f(s, a) + mul8_7f(d, 255u - a);
    v = v > 255u ? 255u : v;
    out |= v << (8 * c);
  }
  return out;
}

// span[(slot * nr + ri) * 4 + {0,1,2,3}] = first stroke edge, stroke edges, first fill edge, fill edges
__global__ void build_ball_sprites_kernel(uint32_t *atlas, RenderLayout RL, int n_balls, const uint32_t *fill8,
                                          const uint32_t *stroke8, const GridEdge *edges, const int32_t *span,
                                          int32_t *overflow) {
  const int slot = blockIdx.y, ri = blockIdx.z;
  if (slot >= n_balls) return;
  const int r = RL.r_min + ri, sT = RL.s_thick[slot];
  const int sz = 2 * (r + sT);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= sz * sz) return;
  const int px = p % sz, py = p / sz;
  const int32_t *sp = span + ((size_t)slot * gridDim.z + ri) * 4;
  uint32_t px8 = 0u;                                   // zeros + paint(1, 1, 1, 0): transparent (primitives.py:45-46)
  if (sp[1] > 0) {                                     // arc, set_line_width, stroke (primitives.py:47-54)
    const int a = pixel_area(edges + sp[0], sp[1], px, py);
    if (a < 0) { atomicAdd(overflow, 1); return; }
    px8 = paint_through_mask(stroke8[slot], area_to_alpha(a), px8);
  }
  if (sp[3] > 0) {                                     // arc, fill (primitives.py:56-58)
    const int a = pixel_area(edges + sp[2], sp[3], px, py);
    if (a < 0) { atomicAdd(overflow, 1); return; }
    px8 = paint_through_mask(fill8[slot], area_to_alpha(a), px8);
  }
  uint8_t *basep = reinterpret_cast<uint8_t *>(atlas) + (size_t)slot * RL.sprite_slot_stride + (size_t)ri * RL.sprite_stride;
  reinterpret_cast<uint32_t *>(basep)[p] = px8;
}

}  // namespace

#ifdef PPB_CAIRO_HOST_CHECK
// tools/sprite_host_check.cu: the per-pixel evaluation above run on the host, so the logic can be compared with the
// oracle in a container without a GPU (never part of the library)
int host_check_sprite(int r, int s_thick, uint32_t fill8, uint32_t stroke8, uint32_t *out) {
  RenderLayout RL{};
  RL.r_min = RL.r_max = r;
  RL.s_thick[0] = s_thick;
  std::vector<GridEdge> edges;
  std::vector<int32_t> span;
  sprite_polygons(RL, 1, edges, span);
  const int sz = 2 * (r + s_thick);
  for (int p = 0; p < sz * sz; ++p) {
    uint32_t px8 = 0u;
    if (span[1] > 0) {
      const int a = pixel_area(edges.data() + span[0], span[1], p % sz, p / sz);
      if (a < 0) return -1;
      px8 = paint_through_mask(stroke8, area_to_alpha(a), px8);
    }
    if (span[3] > 0) {
      const int a = pixel_area(edges.data() + span[2], span[3], p % sz, p / sz);
      if (a < 0) return -1;
      px8 = paint_through_mask(fill8, area_to_alpha(a), px8);
    }
    out[p] = px8;
  }
  return sz;
}
#endif

cudaError_t launch_build_sprites(uint32_t *atlas, const RenderLayout &RL, int n_balls, const uint32_t *fill8,
                                 const uint32_t *stroke8, cudaStream_t st) {
  const int nr = RL.r_max - RL.r_min + 1;
  int max_sz = 0;
  for (int b = 0; b < n_balls; ++b) {
    const int sz = 2 * (RL.r_max + RL.s_thick[b]);
    if (sz > max_sz) max_sz = sz;
  }
  if (nr <= 0 || max_sz <= 0) return cudaSuccess;
  std::vector<GridEdge> edges;
  std::vector<int32_t> span;
  sprite_polygons(RL, n_balls, edges, span);
  GridEdge *d_edges = nullptr;
  int32_t *d_span = nullptr, *d_over = nullptr;
  cudaError_t err = cudaMalloc((void **)&d_edges, (edges.size() + 1) * sizeof(GridEdge));
  if (err != cudaSuccess) return err;
  if ((err = cudaMalloc((void **)&d_span, span.size() * sizeof(int32_t) + 4)) != cudaSuccess) { cudaFree(d_edges); return err; }
  d_over = d_span + span.size();
  int32_t h_over = 0;
  cudaMemcpyAsync(d_edges, edges.data(), edges.size() * sizeof(GridEdge), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_span, span.data(), span.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
  cudaMemsetAsync(d_over, 0, 4, st);
  dim3 grid((max_sz * max_sz + 127) / 128, n_balls, nr);
  build_ball_sprites_kernel<<<grid, 128, 0, st>>>(atlas, RL, n_balls, fill8, stroke8, d_edges, d_span, d_over);
  err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemcpyAsync(&h_over, d_over, 4, cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);     // the host vectors and the scratch go away below
  cudaFree(d_edges);
  cudaFree(d_span);
  if (err == cudaSuccess && h_over != 0) err = cudaErrorInvalidValue;   // a sprite row met more than kMaxActive edges
  return err;
}

}  // namespace ppb
