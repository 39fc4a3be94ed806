/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the drawing pipeline the reference reaches through
 * pycairo (primitives.py:33-77 get_ball_im / get_rectangle_im, :129-174 get_block_im, :237-257
 * GenericWall.imprint, :306-317 Wall.imprint, :423-445 Ball.imprint, :991-1008 generate_image).
 * Never used on the product path.
 *
 * The arithmetic lives in a third-party dependency that is absent from /root/reference and from this
 * image: libcairo + pixman behind pycairo, version unpinned by the reference (README.md:5-9 only says
 * libcairo2-dev).  This file restates the PUBLISHED algorithm of cairo's image backend as shipped in the
 * 1.12 - 1.16 series (the series current when the reference was written), stage by stage:
 *
 *   cairo_arc            cairo-arc.c           arcs > pi are halved; each piece is split into
 *                                              ceil(angle / max_angle(tolerance / radius)) cubic Beziers with
 *                                              h = 4/3 tan(angle / 4)                      (arc_* below)
 *   fixed point          cairo-fixed-private.h every path coordinate is rounded (half to even) to 24.8
 *   flattening           cairo-spline.c        recursive de Casteljau halving in fixed point until both inner
 *                                              control points are within `tolerance` (0.1) of the chord
 *   stroke               cairo-path-stroke-polygon.c
 *                                              curves: one face (point +- half_width * normal of the spline
 *                                              tangent, rounded to 24.8) per flattened point, the two offset
 *                                              contours closed through the butt caps; lines: a rectangle per
 *                                              segment, a miter wedge per join (limit 10), butt caps
 *   scan conversion      cairo-tor-scan-converter.c
 *                                              15 sample rows per pixel (sampled at the TOP of each sub-row),
 *                                              x in 1/256 px from a floored DDA, exact horizontal coverage of
 *                                              each sub-row, non-zero winding,
 *                                              alpha = (17 * area + 256) >> 9 with full area = 2*256*15
 *   unaligned boxes      cairo-rectangular-scan-converter.c   exact area on the 1/256 grid, round half up
 *   solid opaque source  cairo-image-compositor.c (inplace spans)
 *                                              dst = lerp(pixel, alpha, dst), 8-bit multiplies biased 0x7f
 *   anything else        pixman combine_over_u with an a8 mask: (src IN mask) OVER dst, multiplies biased 0x80
 *
 * PARITY UNPINNED: nothing here could be compared with libcairo's output (no pycairo wheel, no libcairo
 * headers or binary in the image, no golden images in the reference).  What IS pinned by tests: the stage
 * parameters above as known-answer tests (segment counts, flattened vertex counts, the alpha map of full /
 * half / empty pixels, lerp and OVER identities) and agreement of the product's device rasteriser with this
 * file.  The older exact-area model (render_oracle.c) is kept so the difference between the two is a
 * measured number (tests/test_cairo_oracle.py), not an estimate.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t fx_t;                 /* cairo_fixed_t, 24.8 */
typedef struct { fx_t x, y; } pt_t;
#define FX_ONE 256
#define GRID_Y 15
#define GRID_X 256
#define GRID_XY (2 * GRID_X * GRID_Y)
#define CAIRO_TOLERANCE 0.1            /* cairo's default gstate tolerance */

/* _cairo_fixed_from_double: the magic-number trick rounds to nearest, ties to even */
static fx_t fx_from_double(double d) { return (fx_t)nearbyint(d * 256.0); }
static double fx_to_double(fx_t f) { return (double)f / 256.0; }

/* ------------------------------------------------------------------ polygon (cairo-polygon.c) */
typedef struct { fx_t x1, y1, x2, y2; int dir; } edge_t;        /* y1 < y2 */
typedef struct { edge_t *e; int n, cap; } poly_t;

static void poly_init(poly_t *p) { p->e = NULL; p->n = p->cap = 0; }
static void poly_free(poly_t *p) { free(p->e); p->e = NULL; p->n = p->cap = 0; }
static void poly_add_edge(poly_t *p, pt_t a, pt_t b, int dir) {
  if (a.y == b.y) return;                                       /* horizontal edges carry no coverage */
  if (a.y > b.y) { pt_t t = a; a = b; b = t; dir = -dir; }
  if (p->n == p->cap) {
    p->cap = p->cap ? 2 * p->cap : 64;
    p->e = (edge_t *)realloc(p->e, sizeof(edge_t) * (size_t)p->cap);
  }
  edge_t e = {a.x, a.y, b.x, b.y, dir};
  p->e[p->n++] = e;
}
static void poly_add_closed(poly_t *p, const pt_t *v, int n, int dir) {
  for (int i = 0; i < n; ++i) poly_add_edge(p, v[i], v[(i + 1) % n], dir);
}

/* ------------------------------------------------------------------ splines (cairo-spline.c) */
typedef void (*add_point_fn)(void *closure, pt_t p, fx_t tdx, fx_t tdy);
typedef struct { pt_t a, b, c, d; } knots_t;
typedef struct { add_point_fn fn; void *closure; pt_t last; } spline_t;

static double spline_error_squared(const knots_t *k) {
  double bdx = fx_to_double(k->b.x - k->a.x), bdy = fx_to_double(k->b.y - k->a.y);
  double cdx = fx_to_double(k->c.x - k->a.x), cdy = fx_to_double(k->c.y - k->a.y);
  if (k->a.x != k->d.x || k->a.y != k->d.y) {
    double dx = fx_to_double(k->d.x - k->a.x), dy = fx_to_double(k->d.y - k->a.y);
    double v = dx * dx + dy * dy, u = bdx * dx + bdy * dy;
    if (u <= 0) { } else if (u >= v) { bdx -= dx; bdy -= dy; } else { bdx -= u / v * dx; bdy -= u / v * dy; }
    u = cdx * dx + cdy * dy;
    if (u <= 0) { } else if (u >= v) { cdx -= dx; cdy -= dy; } else { cdx -= u / v * dx; cdy -= u / v * dy; }
  }
  double berr = bdx * bdx + bdy * bdy, cerr = cdx * cdx + cdy * cdy;
  return berr > cerr ? berr : cerr;
}
static pt_t lerp_half(pt_t a, pt_t b) { pt_t r = {a.x + ((b.x - a.x) >> 1), a.y + ((b.y - a.y) >> 1)}; return r; }
static void spline_add_point(spline_t *s, pt_t p, pt_t knot) {
  if (s->last.x == p.x && s->last.y == p.y) return;
  s->last = p;
  s->fn(s->closure, p, knot.x - p.x, knot.y - p.y);
}
static void spline_decompose_into(knots_t *s1, double tol2, spline_t *res) {
  if (spline_error_squared(s1) < tol2) { spline_add_point(res, s1->a, s1->b); return; }
  knots_t s2;
  pt_t ab = lerp_half(s1->a, s1->b), bc = lerp_half(s1->b, s1->c), cd = lerp_half(s1->c, s1->d);
  pt_t abbc = lerp_half(ab, bc), bccd = lerp_half(bc, cd), fin = lerp_half(abbc, bccd);
  s2.a = fin; s2.b = bccd; s2.c = cd; s2.d = s1->d;
  s1->b = ab; s1->c = abbc; s1->d = fin;
  spline_decompose_into(s1, tol2, res);
  spline_decompose_into(&s2, tol2, res);
}
/* _cairo_spline_decompose: every flattened point but the first, each with the tangent cairo hands the
 * stroker; the end point comes with the spline's final slope */
static void spline_decompose(pt_t a, pt_t b, pt_t c, pt_t d, double tol, add_point_fn fn, void *closure) {
  spline_t s = {fn, closure, a};
  knots_t k = {a, b, c, d};
  pt_t fs_from = c;
  if (c.x == d.x && c.y == d.y) fs_from = (b.x == d.x && b.y == d.y) ? a : b;
  spline_decompose_into(&k, tol * tol, &s);
  fn(closure, d, d.x - fs_from.x, d.y - fs_from.y);
}

/* ------------------------------------------------------------------ arcs (cairo-arc.c) */
static double arc_error_normalized(double angle) { return 2.0 / 27.0 * pow(sin(angle / 4), 6) / pow(cos(angle / 4), 2); }
static double arc_max_angle_for_tolerance_normalized(double tolerance) {
  static const struct { double angle, error; } table[] = {
      {M_PI / 1.0, 0.0185185185185185036127},  {M_PI / 2.0, 0.000272567143730179811158},
      {M_PI / 3.0, 2.38647043651461047433e-05}, {M_PI / 4.0, 4.2455377443222443279e-06},
      {M_PI / 5.0, 1.11281001494389081528e-06}, {M_PI / 6.0, 3.72662000942734705475e-07},
      {M_PI / 7.0, 1.47783685574284411325e-07}, {M_PI / 8.0, 6.63240432022601149057e-08},
      {M_PI / 9.0, 3.2715520137536980553e-08},  {M_PI / 10.0, 1.73863223499021216974e-08},
      {M_PI / 11.0, 9.81410988043554039085e-09}};
  int n = (int)(sizeof(table) / sizeof(table[0])), i;
  for (i = 0; i < n; ++i)
    if (table[i].error < tolerance) return table[i].angle;
  double angle, error;
  ++i;
  do { angle = M_PI / i++; error = arc_error_normalized(angle); } while (error > tolerance);
  return angle;
}
int cairo_arc_segments_needed(double angle, double radius, double tolerance) {
  return (int)ceil(fabs(angle) / arc_max_angle_for_tolerance_normalized(tolerance / radius));
}
typedef struct { pt_t p[4]; } bez_t;                       /* p[0] = current point */
typedef struct { bez_t *b; int n; pt_t cur; int has_cur; pt_t first; } path_t;
static void path_curve(path_t *P, double x1, double y1, double x2, double y2, double x3, double y3) {
  bez_t z;
  z.p[0] = P->cur;
  z.p[1].x = fx_from_double(x1); z.p[1].y = fx_from_double(y1);
  z.p[2].x = fx_from_double(x2); z.p[2].y = fx_from_double(y2);
  z.p[3].x = fx_from_double(x3); z.p[3].y = fx_from_double(y3);
  P->b[P->n++] = z;
  P->cur = z.p[3];
}
static void arc_segment(path_t *P, double xc, double yc, double radius, double A, double B) {
  double r_sin_A = radius * sin(A), r_cos_A = radius * cos(A);
  double r_sin_B = radius * sin(B), r_cos_B = radius * cos(B);
  double h = 4.0 / 3.0 * tan((B - A) / 4.0);
  path_curve(P, xc + r_cos_A - h * r_sin_A, yc + r_sin_A + h * r_cos_A, xc + r_cos_B + h * r_sin_B,
             yc + r_sin_B - h * r_cos_B, xc + r_cos_B, yc + r_sin_B);
}
static void arc_in_direction(path_t *P, double xc, double yc, double radius, double amin, double amax) {
  if (amax - amin > M_PI) {
    double amid = amin + (amax - amin) / 2.0;
    arc_in_direction(P, xc, yc, radius, amin, amid);
    arc_in_direction(P, xc, yc, radius, amid, amax);
  } else if (amax != amin) {
    int segments = cairo_arc_segments_needed(amax - amin, radius, CAIRO_TOLERANCE);
    double step = (amax - amin) / segments;
    segments -= 1;
    /* cairo >= 1.14 also issues a line_to(start of this piece): a zero-length segment here */
    for (int i = 0; i < segments; ++i, amin += step) arc_segment(P, xc, yc, radius, amin, amin + step);
    arc_segment(P, xc, yc, radius, amin, amax);
  }
}
/* cr.arc(xc, yc, radius, 0, 2*pi) on a context without a current point */
static void path_full_circle(path_t *P, bez_t *storage, double xc, double yc, double radius) {
  P->b = storage; P->n = 0;
  P->cur.x = fx_from_double(xc + radius * cos(0.0)); P->cur.y = fx_from_double(yc + radius * sin(0.0));
  P->first = P->cur; P->has_cur = 1;
  arc_in_direction(P, xc, yc, radius, 0.0, 2 * M_PI);
}

/* ------------------------------------------------------------------ fill: path -> polygon */
typedef struct { pt_t *v; int n, cap; } ptlist_t;
static void ptlist_push(ptlist_t *L, pt_t p) {
  if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 64; L->v = (pt_t *)realloc(L->v, sizeof(pt_t) * (size_t)L->cap); }
  L->v[L->n++] = p;
}
static void fill_add_point(void *closure, pt_t p, fx_t tdx, fx_t tdy) { (void)tdx; (void)tdy; ptlist_push((ptlist_t *)closure, p); }
static void path_fill_to_polygon(const path_t *P, poly_t *poly) {
  ptlist_t L = {NULL, 0, 0};
  ptlist_push(&L, P->first);
  for (int i = 0; i < P->n; ++i)
    spline_decompose(P->b[i].p[0], P->b[i].p[1], P->b[i].p[2], P->b[i].p[3], CAIRO_TOLERANCE, fill_add_point, &L);
  poly_add_closed(poly, L.v, L.n, 1);               /* fill closes the sub-path implicitly */
  free(L.v);
}
int cairo_circle_flatten(double xc, double yc, double radius, int max_pts, int32_t *xy) {
  bez_t st[64]; path_t P; ptlist_t L = {NULL, 0, 0};
  path_full_circle(&P, st, xc, yc, radius);
  ptlist_push(&L, P.first);
  for (int i = 0; i < P.n; ++i)
    spline_decompose(P.b[i].p[0], P.b[i].p[1], P.b[i].p[2], P.b[i].p[3], CAIRO_TOLERANCE, fill_add_point, &L);
  int n = L.n;
  for (int i = 0; i < n && i < max_pts; ++i) { xy[2 * i] = L.v[i].x; xy[2 * i + 1] = L.v[i].y; }
  free(L.v);
  return n;
}

/* ------------------------------------------------------------------ stroke (cairo-path-stroke-polygon.c) */
typedef struct { pt_t ccw, point, cw; double sx, sy; } face_t;
static void compute_face(pt_t p, fx_t dx, fx_t dy, double half_width, face_t *f) {
  double sx = fx_to_double(dx), sy = fx_to_double(dy);
  if (sx == 0.0 && sy == 0.0) { }
  else if (sx == 0.0) { sy = sy > 0 ? 1.0 : -1.0; }
  else if (sy == 0.0) { sx = sx > 0 ? 1.0 : -1.0; }
  else { double m = sqrt(sx * sx + sy * sy); sx /= m; sy /= m; }   /* cairo: hypot(); <= 1 ulp apart, then rounded to 1/256 */
  f->sx = sx; f->sy = sy;
  fx_t ox = fx_from_double(-sy * half_width), oy = fx_from_double(sx * half_width);
  f->point = p;
  f->ccw.x = p.x + ox; f->ccw.y = p.y + oy;
  f->cw.x = p.x - ox; f->cw.y = p.y - oy;
}
typedef struct { ptlist_t ccw, cw; double half_width, cusp_tol; face_t cur; int cusps; } stroker_t;
static void stroke_spline_to(void *closure, pt_t p, fx_t tdx, fx_t tdy) {
  stroker_t *S = (stroker_t *)closure;
  face_t f;
  if ((tdx | tdy) == 0) { S->cusps++; return; }
  compute_face(p, tdx, tdy, S->half_width, &f);
  if (f.sx * S->cur.sx + f.sy * S->cur.sy < S->cusp_tol) S->cusps++;   /* cairo inserts a round fan here */
  ptlist_push(&S->ccw, f.ccw);
  ptlist_push(&S->cw, f.cw);
  S->cur = f;
}
/* stroke of an open path made of curve_to's only (the ball's ring): returns the number of cusp / degenerate
 * tangents met (0 for every circle; the fan cairo would add there is not restated) */
static int path_stroke_curves_to_polygon(const path_t *P, double line_width, poly_t *poly) {
  stroker_t S;
  memset(&S, 0, sizeof S);
  S.half_width = line_width / 2.0;
  S.cusp_tol = 1 - CAIRO_TOLERANCE / S.half_width;
  S.cusp_tol *= S.cusp_tol; S.cusp_tol *= 2; S.cusp_tol -= 1;
  for (int i = 0; i < P->n; ++i) {
    const bez_t *z = &P->b[i];
    face_t f;
    compute_face(z->p[0], z->p[1].x - z->p[0].x, z->p[1].y - z->p[0].y, S.half_width, &f);
    if (i == 0) {
      ptlist_push(&S.ccw, f.ccw);
      ptlist_push(&S.cw, f.cw);
    } else if (f.ccw.x != S.cur.ccw.x || f.ccw.y != S.cur.ccw.y || f.cw.x != S.cur.cw.x || f.cw.y != S.cur.cw.y) {
      /* join between two Beziers of one arc: tangent-continuous, so the faces differ by at most one 1/256
       * step; outer side: the new face point (bevel == miter at this size), inner side: cairo's inner_join
       * passes through the centre point */
      double cross = S.cur.sx * f.sy - S.cur.sy * f.sx;
      if (cross > 0) { ptlist_push(&S.cw, f.cw); ptlist_push(&S.ccw, S.cur.point); ptlist_push(&S.ccw, f.ccw); }   /* cw side is outer */
      else { ptlist_push(&S.ccw, f.ccw); ptlist_push(&S.cw, S.cur.point); ptlist_push(&S.cw, f.cw); }
    }
    S.cur = f;
    spline_decompose(z->p[0], z->p[1], z->p[2], z->p[3], CAIRO_TOLERANCE, stroke_spline_to, &S);
  }
  /* butt caps: the loop ccw[0..n) -> cw[n-1..0] -> ccw[0] */
  for (int i = 0; i + 1 < S.ccw.n; ++i) poly_add_edge(poly, S.ccw.v[i], S.ccw.v[i + 1], 1);
  if (S.ccw.n && S.cw.n) poly_add_edge(poly, S.ccw.v[S.ccw.n - 1], S.cw.v[S.cw.n - 1], 1);
  for (int i = S.cw.n - 1; i > 0; --i) poly_add_edge(poly, S.cw.v[i], S.cw.v[i - 1], 1);
  if (S.ccw.n && S.cw.n) poly_add_edge(poly, S.cw.v[0], S.ccw.v[0], 1);
  free(S.ccw.v); free(S.cw.v);
  return S.cusps;
}
static void add_piece(poly_t *poly, pt_t *v, int n) {     /* orient every piece the same way, then add */
  double a = 0;
  for (int i = 0; i < n; ++i) { int j = (i + 1) % n; a += (double)v[i].x * v[j].y - (double)v[j].x * v[i].y; }
  if (a == 0) return;
  if (a < 0) for (int i = 0; i < n / 2; ++i) { pt_t t = v[i]; v[i] = v[n - 1 - i]; v[n - 1 - i] = t; }
  poly_add_closed(poly, v, n, 1);
}
/* stroke of an open polyline (get_block_im's move_to + 4 line_to, :160-166): miter joins (limit 10), butt caps.
 * The stroked region is the union of one rectangle per segment and one wedge per join -- the same set
 * cairo's two contours enclose under the non-zero rule. */
static void polyline_stroke_to_polygon(const pt_t *p, int n, double line_width, poly_t *poly) {
  double hw = line_width / 2.0, ml = 10.0;
  face_t prev_end; int have_prev = 0;
  pt_t cur = p[0];
  for (int i = 1; i < n; ++i) {
    if (p[i].x == cur.x && p[i].y == cur.y) continue;
    fx_t dx = p[i].x - cur.x, dy = p[i].y - cur.y;
    face_t start, end;
    compute_face(cur, dx, dy, hw, &start);
    end = start;
    end.point = p[i];
    end.ccw.x += dx; end.ccw.y += dy; end.cw.x += dx; end.cw.y += dy;
    pt_t rect[4] = {start.ccw, end.ccw, end.cw, start.cw};
    add_piece(poly, rect, 4);
    if (have_prev) {
      const face_t *in = &prev_end, *out = &start;
      double cross = in->sx * out->sy - in->sy * out->sx;
      int ccw_is_outer = cross < 0;                   /* the path turns away from its ccw side */
      pt_t inpt = ccw_is_outer ? in->ccw : in->cw, outpt = ccw_is_outer ? out->ccw : out->cw;
      if (inpt.x != outpt.x || inpt.y != outpt.y) {
        double in_dot_out = in->sx * out->sx + in->sy * out->sy;
        int done = 0;
        if (2 <= ml * ml * (1 + in_dot_out)) {
          double x1 = fx_to_double(inpt.x), y1 = fx_to_double(inpt.y), dx1 = in->sx, dy1 = in->sy;
          double x2 = fx_to_double(outpt.x), y2 = fx_to_double(outpt.y), dx2 = out->sx, dy2 = out->sy;
          double my = ((x2 - x1) * dy1 * dy2 - y2 * dx2 * dy1 + y1 * dx1 * dy2) / (dx1 * dy2 - dx2 * dy1);
          double mx = fabs(dy1) >= fabs(dy2) ? (my - y1) * dx1 / dy1 + x1 : (my - y2) * dx2 / dy2 + x2;
          double ix = fx_to_double(in->point.x), iy = fx_to_double(in->point.y);
          double fdx1 = x1 - ix, fdy1 = y1 - iy, fdx2 = x2 - ix, fdy2 = y2 - iy, mdx = mx - ix, mdy = my - iy;
          double c1 = fdx1 * mdy - fdy1 * mdx, c2 = fdx2 * mdy - fdy2 * mdx;
          int s1 = (c1 > 0) - (c1 < 0), s2 = (c2 > 0) - (c2 < 0);
          if (s1 != s2) {
            pt_t m = {fx_from_double(mx), fx_from_double(my)};
            pt_t wedge[4] = {in->point, inpt, m, outpt};
            add_piece(poly, wedge, 4);
            done = 1;
          }
        }
        if (!done) { pt_t bevel[3] = {in->point, inpt, outpt}; add_piece(poly, bevel, 3); }
      }
    }
    prev_end = end; have_prev = 1;
    cur = p[i];
  }
}

/* ------------------------------------------------------------------ cairo-tor-scan-converter.c */
static int64_t floor_div(int64_t a, int64_t b) {           /* floored_divrem / floored_muldivrem */
  int64_t q = a / b, r = a % b;
  if (r != 0 && ((r < 0) != (b < 0))) --q;
  return q;
}
static int32_t input_to_grid_y(fx_t in) { return (int32_t)(((int64_t)GRID_Y * in + (1 << 7)) >> 8); }
typedef struct { int32_t ytop, ybot, y1, dy, x1, dx; int dir; } tedge_t;
#define GRID_AREA_TO_ALPHA(c) ((((c) + ((c) << 4)) + 256) >> 9)
/* a8 coverage of `poly` (non-zero winding) on a w x h surface whose pixel (0,0) covers [0,1)x[0,1).
 * glitter_scan_converter_render: a pixel row in which no edge starts or ends and no two edges cross is
 * covered analytically, one trapezoid per edge (cell_list_render_edge, with the crossing heights still
 * counted in whole sub-rows); every other row is sampled on its 15 sub-rows. */
typedef struct { int32_t x, xb; int dir; } xdir_t2;
static int cmp_xdir2(const void *a, const void *b) {
  const xdir_t2 *p = (const xdir_t2 *)a, *q = (const xdir_t2 *)b;
  if (p->x != q->x) return (p->x > q->x) - (p->x < q->x);
  return (p->xb > q->xb) - (p->xb < q->xb);
}
static int32_t tedge_x(const tedge_t *e, int32_t y) {
  if (e->dx == 0) return e->x1;
  return e->x1 + (int32_t)floor_div((int64_t)(y - e->y1) * e->dx, e->dy);
}
static void cell_add(int32_t *cov_h, int32_t *unc, int w, int ix, int32_t dcov, int32_t dunc) {
  /* blit: cells left of the surface only pass their covered height on ("skip cells to the left of the clip
   * region"), cells right of it are never read */
  if (ix < 0) { cov_h[0] += dcov; return; }
  if (ix > w) ix = w;
  cov_h[ix] += dcov; unc[ix] += dunc;
}
static void render_edge_full_row(int32_t *cov_h, int32_t *unc, int w, int32_t x1q, int32_t x2q, int sign) {
  int ix1 = x1q >> 8, ix2 = x2q >> 8;
  int32_t fx1 = x1q & 255, fx2 = x2q & 255;
  if (ix1 == ix2) { cell_add(cov_h, unc, w, ix1, sign * GRID_Y, sign * (fx1 + fx2) * GRID_Y); return; }
  int32_t dx = x2q - x1q, y1, y2, dy;
  if (dx >= 0) { y1 = 0; y2 = GRID_Y; }
  else {
    int t = ix1; ix1 = ix2; ix2 = t;
    int32_t f = fx1; fx1 = fx2; fx2 = f;
    dx = -dx; sign = -sign; y1 = GRID_Y; y2 = 0;
  }
  dy = y2 - y1;
  int64_t num = (int64_t)(GRID_X - fx1) * dy;
  int32_t yq = (int32_t)floor_div(num, dx), yr = (int32_t)(num - (int64_t)yq * dx);
  cell_add(cov_h, unc, w, ix1, sign * yq, sign * yq * (GRID_X + fx1));
  yq += y1;
  if (ix1 + 1 < ix2) {
    int64_t fnum = (int64_t)GRID_X * dy;
    int32_t fq = (int32_t)floor_div(fnum, dx), fr = (int32_t)(fnum - (int64_t)fq * dx);
    ++ix1;
    do {
      int32_t y_skip = fq;
      yr += fr;
      if (yr >= dx) { ++y_skip; yr -= dx; }
      yq += y_skip;
      y_skip *= sign;
      cell_add(cov_h, unc, w, ix1, y_skip, y_skip * GRID_X);
      ++ix1;
    } while (ix1 != ix2);
  }
  cell_add(cov_h, unc, w, ix2, sign * (y2 - yq), sign * (y2 - yq) * fx2);
}
int cairo_tor_rows_full = 0, cairo_tor_rows_sub = 0;         /* statistics for the tests */
static void tor_scan_convert(const poly_t *poly, int w, int h, uint8_t *a8) {
  int ne = 0;
  tedge_t *E = (tedge_t *)malloc(sizeof(tedge_t) * (size_t)(poly->n > 0 ? poly->n : 1));
  int32_t ymin = 0, ymax = h * GRID_Y;
  for (int i = 0; i < poly->n; ++i) {
    const edge_t *e = &poly->e[i];
    tedge_t t;
    int32_t top = input_to_grid_y(e->y1), bot = input_to_grid_y(e->y2);     /* an edge spans its whole line */
    if (top >= bot) continue;
    if (top >= ymax || bot <= ymin) continue;
    t.y1 = top;
    t.ytop = top >= ymin ? top : ymin;
    t.ybot = bot <= ymax ? bot : ymax;
    t.x1 = e->x1; t.dx = e->x2 - e->x1; t.dy = bot - top; t.dir = e->dir;
    E[ne++] = t;
  }
  int32_t *cov_h = (int32_t *)malloc(sizeof(int32_t) * (size_t)(w + 2));
  int32_t *unc = (int32_t *)malloc(sizeof(int32_t) * (size_t)(w + 2));
  xdir_t2 *act = (xdir_t2 *)malloc(sizeof(xdir_t2) * (size_t)(ne > 0 ? ne : 1));
  int32_t xmax = w * GRID_X;
  for (int row = 0; row < h; ++row) {
    int32_t y0 = row * GRID_Y;
    memset(cov_h, 0, sizeof(int32_t) * (size_t)(w + 2));
    memset(unc, 0, sizeof(int32_t) * (size_t)(w + 2));
    /* can this row be stepped as a whole?  (y_buckets[row] empty, min_height >= GRID_Y, no crossing) */
    int bucket = 0, na = 0, full = 1;
    for (int i = 0; i < ne; ++i) {
      if (E[i].ytop >= y0 && E[i].ytop < y0 + GRID_Y) bucket = 1;
      if (E[i].ytop <= y0 && y0 < E[i].ybot) {
        if (E[i].ybot - y0 < GRID_Y) full = 0;
        act[na].x = tedge_x(&E[i], y0);
        act[na].xb = tedge_x(&E[i], y0 + GRID_Y);
        act[na].dir = E[i].dir;
        ++na;
      }
    }
    if (bucket) full = 0;
    if (na == 0 && !bucket) { memset(a8 + (size_t)row * w, 0, (size_t)w); continue; }
    if (full) {
      qsort(act, (size_t)na, sizeof(xdir_t2), cmp_xdir2);
      int64_t prev = INT64_MIN;
      for (int i = 0; i < na; ++i) { if (act[i].xb <= prev) { full = 0; break; } prev = act[i].xb; }
    }
    if (full) {
      ++cairo_tor_rows_full;
      int i = 0;
      while (i < na) {                                  /* apply_nonzero_fill_rule_and_step_edges */
        int left = i, winding = act[i].dir;
        while (1) {
          ++i;
          if (i >= na) break;
          winding += act[i].dir;
          if (winding == 0 && (i + 1 >= na || act[i + 1].x != act[i].x)) break;
        }
        if (i >= na) break;                             /* unbalanced: cannot happen for closed polygons */
        render_edge_full_row(cov_h, unc, w, act[left].x, act[left].xb, +1);
        render_edge_full_row(cov_h, unc, w, act[i].x, act[i].xb, -1);
        ++i;
      }
    } else {
      ++cairo_tor_rows_sub;
      for (int sub = 0; sub < GRID_Y; ++sub) {
        int32_t y = y0 + sub;
        na = 0;
        for (int i = 0; i < ne; ++i)
          if (E[i].ytop <= y && y < E[i].ybot) {
            act[na].x = tedge_x(&E[i], y);
            act[na].xb = 0;
            act[na].dir = E[i].dir;
            ++na;
          }
        qsort(act, (size_t)na, sizeof(xdir_t2), cmp_xdir2);
        int winding = 0; int32_t xstart = 0;
        for (int i = 0; i < na; ++i) {
          if (winding == 0) xstart = act[i].x;
          winding += act[i].dir;
          if (winding == 0) {                            /* cell_list_add_subspan(xstart, x) */
            int32_t x1 = xstart, x2 = act[i].x;
            if (x1 < 0) x1 = 0;
            if (x2 > xmax) x2 = xmax;
            if (x1 >= x2) continue;
            int ix1 = x1 >> 8, fx1 = x1 & 255, ix2 = x2 >> 8, fx2 = x2 & 255;
            if (ix1 != ix2) {
              unc[ix1] += 2 * fx1; cov_h[ix1] += 1;
              unc[ix2] -= 2 * fx2; cov_h[ix2] -= 1;
            } else {
              unc[ix1] += 2 * (fx1 - fx2);
            }
          }
        }
      }
    }
    int32_t cover = 0;
    for (int x = 0; x < w; ++x) {
      cover += cov_h[x] * GRID_X * 2;
      int32_t area = cover - unc[x];
      if (area < 0) area = 0;
      if (area > GRID_XY) area = GRID_XY;
      a8[(size_t)row * w + x] = (uint8_t)GRID_AREA_TO_ALPHA(area);
    }
  }
  free(E); free(cov_h); free(unc); free(act);
}

/* ------------------------------------------------------------------ compositing */
static inline uint32_t mul8_7f(uint32_t a, uint32_t b) { uint32_t t = a * b + 0x7f; return ((t >> 8) + t) >> 8; }
static inline uint8_t lerp8(uint32_t src, uint32_t a, uint32_t dst) {         /* lerp8x4, one channel */
  uint32_t t = mul8_7f(src, a) + mul8_7f(dst, 255u - a);
  return (uint8_t)(t > 255u ? 255u : t);
}
static inline uint32_t mul_un8(uint32_t a, uint32_t b) { uint32_t t = a * b + 0x80u; return ((t >> 8) + t) >> 8; }
static inline void over_masked(const uint8_t *src, uint32_t m, uint8_t *dst) {   /* pixman combine_over_u + a8 mask */
  uint32_t s[4];
  for (int k = 0; k < 4; ++k) s[k] = mul_un8(src[k], m);
  uint32_t ia = 255u - s[3];
  for (int k = 0; k < 4; ++k) { uint32_t v = s[k] + mul_un8(dst[k], ia); dst[k] = (uint8_t)(v > 255u ? 255u : v); }
}
static inline uint8_t col8(double v) {             /* _cairo_color_double_to_short >> 8 */
  if (v < 0) v = 0;
  if (v > 1) v = 1;
  return (uint8_t)(((uint32_t)(v * 65535.0 + 0.5)) >> 8);
}
static void premul(const float *rgba, uint8_t *p) {
  double a = rgba[3];
  p[0] = col8(rgba[0] * a); p[1] = col8(rgba[1] * a); p[2] = col8(rgba[2] * a); p[3] = col8(a);
}
/* paint a solid colour through an a8 mask the way the image compositor's span renderers do */
static void composite_solid_mask(const float *rgba, const uint8_t *a8, int n, uint8_t *dst) {
  uint8_t px[4];
  premul(rgba, px);
  int opaque = px[3] == 255;
  for (int i = 0; i < n; ++i) {
    uint32_t a = a8[i];
    if (!a) continue;
    uint8_t *d = dst + 4 * (size_t)i;
    if (opaque) {                                                  /* _fill_xrgb32_lerp_opaque_spans */
      if (a == 255) memcpy(d, px, 4);
      else for (int k = 0; k < 4; ++k) d[k] = lerp8(px[k], a, d[k]);
    } else {
      over_masked(px, a, d);
    }
  }
}

/* ------------------------------------------------------------------ sprites */
void cairo_disc_mask(double cx, double cy, double r, int w, int h, uint8_t *a8) {
  bez_t st[64]; path_t P; poly_t poly;
  poly_init(&poly);
  path_full_circle(&P, st, cx, cy, r);
  path_fill_to_polygon(&P, &poly);
  tor_scan_convert(&poly, w, h, a8);
  poly_free(&poly);
}
int cairo_ring_mask(double cx, double cy, double r, double line_width, int w, int h, uint8_t *a8) {
  bez_t st[64]; path_t P; poly_t poly;
  poly_init(&poly);
  path_full_circle(&P, st, cx, cy, r);
  int cusps = path_stroke_curves_to_polygon(&P, line_width, &poly);
  tor_scan_convert(&poly, w, h, a8);
  poly_free(&poly);
  return cusps;
}
/* get_ball_im, primitives.py:33-61 */
void cairo_ball_sprite(int r, int s_thick, const float *fill, const float *stroke, uint8_t *out) {
  int sz = 2 * (r + s_thick);
  double c = r + s_thick;
  uint8_t *a8 = (uint8_t *)malloc((size_t)sz * sz + 1);
  memset(out, 0, (size_t)sz * sz * 4);                       /* zeros + paint(1,1,1,0): transparent */
  if (s_thick > 0 && r > 0) {
    cairo_ring_mask(c, c, r, s_thick, sz, sz, a8);           /* arc + set_line_width + stroke, :47-54 */
    composite_solid_mask(stroke, a8, sz * sz, out);
  }
  if (r > 0) {
    cairo_disc_mask(c, c, r, sz, sz, a8);                      /* arc + fill, :56-58 */
    composite_solid_mask(fill, a8, sz * sz, out);
  }
  free(a8);
}

/* mask of a convex quad filled on a w x h surface: boxes for rectilinear paths, the tor converter otherwise */
static int quad_is_rectilinear(const pt_t *q) {
  for (int i = 0; i < 4; ++i) { int j = (i + 1) % 4; if (q[i].x != q[j].x && q[i].y != q[j].y) return 0; }
  return 1;
}
static void box_mask(fx_t x0, fx_t y0, fx_t x1, fx_t y1, int w, int h, uint8_t *a8) {   /* rectangular scan converter */
  memset(a8, 0, (size_t)w * h);
  if (x0 > x1) { fx_t t = x0; x0 = x1; x1 = t; }
  if (y0 > y1) { fx_t t = y0; y0 = y1; y1 = t; }
  for (int py = 0; py < h; ++py) {
    int64_t ya = (int64_t)py * 256, yb = ya + 256;
    int64_t cy = (y1 < yb ? y1 : yb) - (y0 > ya ? y0 : ya);
    if (cy <= 0) continue;
    for (int px = 0; px < w; ++px) {
      int64_t xa = (int64_t)px * 256, xb = xa + 256;
      int64_t cx = (x1 < xb ? x1 : xb) - (x0 > xa ? x0 : xa);
      if (cx <= 0) continue;
      a8[(size_t)py * w + px] = (uint8_t)((cx * cy * 255 + 32768) / 65536);
    }
  }
}
void cairo_quad_mask(const double *qd, int w, int h, uint8_t *a8) {
  pt_t q[4];
  for (int i = 0; i < 4; ++i) { q[i].x = fx_from_double(qd[2 * i]); q[i].y = fx_from_double(qd[2 * i + 1]); }
  if (quad_is_rectilinear(q)) {
    fx_t x0 = q[0].x, x1 = q[0].x, y0 = q[0].y, y1 = q[0].y;
    for (int i = 1; i < 4; ++i) {
      if (q[i].x < x0) x0 = q[i].x;
      if (q[i].x > x1) x1 = q[i].x;
      if (q[i].y < y0) y0 = q[i].y;
      if (q[i].y > y1) y1 = q[i].y;
    }
    box_mask(x0, y0, x1, y1, w, h, a8);
    return;
  }
  poly_t poly;
  poly_init(&poly);
  poly_add_closed(&poly, q, 4, 1);
  tor_scan_convert(&poly, w, h, a8);
  poly_free(&poly);
}
static double py2_round(double x) { return round(x); }       /* Python 2 round(): half away from zero */

/* get_block_im, primitives.py:129-174, for the wall whose world-space vertices are v[4]: the sprite (premultiplied
 * pixels, xs x ys) the reference pre-renders with a 2-px stroke and a fill of the same colour. */
int cairo_rotated_wall_model = 0;   /* 0: cairo; 1: exact-area solid colour for non-rectilinear quads (tests only) */
double oracle_quad_pixel_area(const double *q, double px, double py);   /* render_oracle.c */
int cairo_wall_sprite_solid = 0;     /* measurement switch: treat the block sprite as solid colour (tests only) */
static uint8_t *block_sprite(const double *v, const float *rgba, int *xs, int *ys) {
  double p[8], mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
  for (int i = 0; i < 4; ++i) { p[2 * i] = v[2 * i] - v[0]; p[2 * i + 1] = v[2 * i + 1] - v[1]; }
  for (int i = 0; i < 4; ++i) { if (p[2 * i] < mnx) mnx = p[2 * i]; if (p[2 * i + 1] < mny) mny = p[2 * i + 1]; }
  for (int i = 0; i < 4; ++i) { p[2 * i] -= mnx; p[2 * i + 1] -= mny; }
  for (int i = 0; i < 4; ++i) { if (p[2 * i] > mxx) mxx = p[2 * i]; if (p[2 * i + 1] > mxy) mxy = p[2 * i + 1]; }
  int w = (int)ceil(mxx), h = (int)ceil(mxy);
  *xs = w; *ys = h;
  if (w <= 0 || h <= 0) return NULL;
  uint8_t *im = (uint8_t *)calloc((size_t)w * h, 4), *a8 = (uint8_t *)malloc((size_t)w * h);
  pt_t q[5];
  for (int i = 0; i < 4; ++i) { q[i].x = fx_from_double(p[2 * i]); q[i].y = fx_from_double(p[2 * i + 1]); }
  q[4] = q[0];
  if (cairo_wall_sprite_solid) {
    uint8_t px[4];
    premul(rgba, px);
    for (int i = 0; i < w * h; ++i) memcpy(im + 4 * (size_t)i, px, 4);
    free(a8);
    return im;
  }
  poly_t poly;
  poly_init(&poly);
  polyline_stroke_to_polygon(q, 5, 2.0, &poly);               /* set_line_width(sThick = 2); stroke, :165-167 */
  tor_scan_convert(&poly, w, h, a8);
  composite_solid_mask(rgba, a8, w * h, im);
  poly_free(&poly);
  cairo_quad_mask(p, w, h, a8);                                /* fill, :168-174 */
  composite_solid_mask(rgba, a8, w * h, im);
  free(a8);
  return im;
}
void cairo_block_sprite(const double *v, const float *rgba, int max_w, int max_h, int32_t *wh, uint8_t *out) {
  int w, h;
  uint8_t *im = block_sprite(v, rgba, &w, &h);
  wh[0] = w; wh[1] = h;
  if (im && w <= max_w && h <= max_h) memcpy(out, im, (size_t)w * h * 4);
  free(im);
}

/*
 * World.generate_image for n_worlds worlds; the arguments of oracle_render_frames (render_oracle.c).
 */
int cairo_render_frames(int n_worlds, int n_balls, int n_walls, const int32_t *order, const double *wall,
                        const int32_t *wall_kind, const float *wall_rgba, const double *pos, const double *rad,
                        const int32_t *s_thick, const float *fill, const float *stroke, int W, int H, uint8_t *out,
                        int n_threads) {
  int rmin = 1 << 30, rmax = -1;
  for (size_t i = 0; i < (size_t)n_worlds * n_balls; ++i)
    if (rad[i] >= 0) {
      int r = (int)rad[i];
      if (r < rmin) rmin = r;
      if (r > rmax) rmax = r;
    }
  int nr = rmax >= rmin ? rmax - rmin + 1 : 0;
  uint8_t **spr = (uint8_t **)calloc((size_t)(nr > 0 ? nr : 1) * (n_balls > 0 ? n_balls : 1), sizeof(uint8_t *));
  for (size_t i = 0; i < (size_t)n_worlds * n_balls; ++i)
    if (rad[i] >= 0) {
      int b = (int)(i % n_balls), r = (int)rad[i];
      uint8_t **slot = &spr[(size_t)b * nr + (r - rmin)];
      if (!*slot) {
        int sz = 2 * (r + s_thick[b]);
        *slot = (uint8_t *)malloc((size_t)sz * sz * 4 + 4);
        cairo_ball_sprite(r, s_thick[b], fill + 4 * b, stroke + 4 * b, *slot);
      }
    }
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(static) num_threads(n_threads)
  for (int w = 0; w < n_worlds; ++w) {
    uint8_t *im = out + (size_t)w * W * H * 4;
    uint8_t *a8 = (uint8_t *)malloc((size_t)W * H);
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
        if (wall_kind[q] == 1) {
          /* Wall.imprint, :306-317: the sprite of get_rectangle_im is a solid opaque/translucent rectangle;
           * source = that sprite pasted at the truncated float bounds, mask = cr.rectangle(x, y, sx, sy) */
          int bx0 = (int)x0, by0 = (int)y0, bx1 = (int)x1, by1 = (int)y1;
          uint8_t px[4];
          premul(wall_rgba + 4 * q, px);
          box_mask(fx_from_double(x0), fx_from_double(y0), fx_from_double(x1), fx_from_double(y1), W, H, a8);
          for (int py = by0 > 0 ? by0 : 0; py < (by1 < H ? by1 : H); ++py)
            for (int pxx = bx0 > 0 ? bx0 : 0; pxx < (bx1 < W ? bx1 : W); ++pxx) {
              uint32_t m = a8[(size_t)py * W + pxx];
              if (m) over_masked(px, m, im + ((size_t)py * W + pxx) * 4);
            }
        } else {
          /* GenericWall.imprint, :237-257: source = the block sprite pasted at the rounded top-left,
           * mask = fill of the wall quad */
          int sw, sh;
          int X0 = (int)py2_round(x0), Y0 = (int)py2_round(y0);
          pt_t qf[4];
          for (int i = 0; i < 4; ++i) { qf[i].x = fx_from_double(v[2 * i]); qf[i].y = fx_from_double(v[2 * i + 1]); }
          if (cairo_rotated_wall_model == 1 && !quad_is_rectilinear(qf)) {
            /* the engine's documented stand-in for rotated quads: solid colour through the exact pixel-quad area,
             * inside the rectangle the sprite is pasted into (render_oracle.c) */
            int bx1 = X0 + (int)ceil(x1 - x0), by1 = Y0 + (int)ceil(y1 - y0);
            uint8_t px[4];
            premul(wall_rgba + 4 * q, px);
            for (int py = Y0 > 0 ? Y0 : 0; py < (by1 < H ? by1 : H); ++py)
              for (int pxx = X0 > 0 ? X0 : 0; pxx < (bx1 < W ? bx1 : W); ++pxx) {
                uint32_t m = (uint32_t)floor(255.0 * oracle_quad_pixel_area(v, pxx, py) + 0.5);
                if (m) over_masked(px, m, im + ((size_t)py * W + pxx) * 4);
              }
            continue;
          }
          uint8_t *sp = block_sprite(v, wall_rgba + 4 * q, &sw, &sh);
          cairo_quad_mask(v, W, H, a8);
          if (sp)
            for (int sy = 0; sy < sh; ++sy) {
              int py = Y0 + sy;
              if (py < 0 || py >= H) continue;
              for (int sx = 0; sx < sw; ++sx) {
                int pxx = X0 + sx;
                if (pxx < 0 || pxx >= W) continue;
                uint32_t m = a8[(size_t)py * W + pxx];
                if (m) over_masked(sp + ((size_t)sy * sw + sx) * 4, m, im + ((size_t)py * W + pxx) * 4);
              }
            }
          free(sp);
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
            int pxx = X + sx;
            if (pxx < 0 || pxx >= W) continue;
            over_masked(s + ((size_t)sy * sz + sx) * 4, 255, im + ((size_t)py * W + pxx) * 4);   /* boxes: no mask */
          }
        }
      }
    }
    free(a8);
  }
  for (size_t i = 0; i < (size_t)(nr > 0 ? nr : 1) * (n_balls > 0 ? n_balls : 1); ++i) free(spr[i]);
  free(spr);
  return 0;
}
