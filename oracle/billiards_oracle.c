/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the pyphy-engine
 * simulate-and-render path, in plain C / IEEE fp64.
 *
 * Nothing on the product path may include, link or call this file; it is used
 * by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs, and only as the checker.
 *
 * Pinning: tests/test_oracle_golden.py checks this restatement against
 * trajectories, event lists and failure outcomes produced by the UNMODIFIED
 * reference run in the build container through oracle/ref_shim.py (fixtures in
 * tests/golden/, generator oracle/make_golden.py), and against the geometry
 * known-answer values printed by the reference's unit_tests.py.  Physics parity
 * is therefore pinned.  Frame parity is UNPINNED: the reference renders through
 * pycairo/libcairo (absent from the reference tree, version not stated,
 * absent from this image), so render_frame() below restates the Cairo drawing
 * semantics without a runnable ground truth.
 *
 * Every arithmetic expression follows the reference's order of operations so
 * that, compiled with -ffp-contract=off on x86-64/glibc (the same libm numpy
 * and CPython call), results are bit-identical with the Python reference.
 * Citations are file:line in the reference tree.
 *
 * trig_mode 0: libm acos/asin/sin  (what the reference calls: math.acos,
 *              math.asin, np.sin -- geometry.py:174,443-448)
 * trig_mode 1: the portable, FMA-free implementations in
 *              pyphy_engine_b200/csrc/portable_trig.h that the fp64 CUDA kernels
 *              also use, so device and oracle can be compared bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../pyphy_engine_b200/csrc/portable_trig.h"

#define TOL 1e-08 /* geometry.py:7 */

/* per-world outcome codes; the product library uses the same numbering
 * (include/pyphy_b200.h) so that failures can be compared */
enum {
  ST_OK = 0,
  ST_ASSERT_TCOL_NEG = 1,   /* primitives.py:610 */
  ST_ASSERT_OVERLAP = 2,    /* geometry.py:413 */
  ST_ASSERT_THETA = 3,      /* geometry.py:436-438 */
  ST_ASSERT_THETA_A = 4,    /* geometry.py:447 */
  ST_ASSERT_INTPOINT = 5,   /* dynamics.py:34 */
  ST_TRAP_AMISS = 6,        /* dynamics.py:44-45 */
  ST_MATH_DOMAIN = 7,       /* math.acos / math.asin ValueError, geometry.py:174,444 */
  ST_ASSERT_TCOL_EQ = 8,    /* primitives.py:664 */
  ST_ITER_LIMIT = 9         /* guard shared with the engine (ppb_config.max_substeps) */
};

typedef struct { double x, y; } pt_t;
typedef struct { pt_t st, en; double a, b, c; } line_t;

static int g_trig_mode = 0;
static inline double o_sin(double x)  { return g_trig_mode ? ptrig_sin(x)  : sin(x); }
static inline double o_asin(double x) { return g_trig_mode ? ptrig_asin(x) : asin(x); }
static inline double o_acos(double x) { return g_trig_mode ? ptrig_acos(x) : acos(x); }

/* ---- Point, geometry.py:9-175 ------------------------------------------- */
static inline pt_t P(double x, double y) { pt_t p = {x, y}; return p; }
static inline double pt_mag(pt_t p) { return sqrt(p.x * p.x + p.y * p.y); }       /* :31 */
static inline pt_t pt_scale(pt_t p, double s) { return P(s * p.x, s * p.y); }     /* :34 */
static inline pt_t pt_unit(pt_t p) {                                              /* :48 */
  double m = pt_mag(p);
  if (m > 0) return pt_scale(p, 1.0 / m);
  return p;
}
static inline double pt_dot(pt_t a, pt_t b) { return a.x * b.x + a.y * b.y; }     /* :54 */
static inline pt_t pt_add(pt_t a, pt_t b) { return P(a.x + b.x, a.y + b.y); }     /* :84 */
static inline pt_t pt_sub(pt_t a, pt_t b) { return P(a.x - b.x, a.y - b.y); }     /* :96 */
static inline pt_t pt_mul(pt_t a, double s) { return P(a.x * s, a.y * s); }       /* :108 */
static inline pt_t pt_project(pt_t self, pt_t other) {                            /* :58 */
  pt_t u = pt_unit(other);
  double pm = pt_dot(self, u);
  return pt_scale(u, pm);
}
static inline double pt_cosine(pt_t a, pt_t b) {                                  /* :66 */
  return pt_dot(pt_unit(a), pt_unit(b));
}
static inline pt_t pt_reflect_normal(pt_t nrml, pt_t other) {                     /* :74 */
  pt_t s = pt_unit(nrml);
  pt_t prll = pt_project(other, s);
  pt_t orth = pt_sub(other, prll);
  prll = pt_scale(prll, -1);
  return pt_add(prll, orth);
}
static inline double pt_distance(pt_t a, pt_t b) { return pt_mag(pt_sub(a, b)); } /* :136 */

/* get_angle_between, geometry.py:166-175; *err set on math domain error */
static double pt_angle_between(pt_t a, pt_t b, int *err) {
  double c = pt_dot(pt_unit(a), pt_unit(b));
  if (c < 1 + 1e-6) c = (1 < c) ? 1 : c;      /* min(1, cosTheta) */
  if (!(c >= -1.0 && c <= 1.0)) { *err = ST_MATH_DOMAIN; return 0.0; }
  return o_acos(c);
}

/* ---- Line, geometry.py:191-357 ------------------------------------------ */
static line_t line_make(pt_t p1, pt_t p2) {                                       /* :203 */
  line_t l;
  l.st = p1; l.en = p2;
  l.a = -(p2.y - p1.y);
  l.b = p2.x - p1.x;
  l.c = p1.x * p2.y - p1.y * p2.x;
  double am = fabs(l.a);
  if (am > TOL) { l.a = l.a / am; l.b = l.b / am; l.c = l.c / am; }
  else l.a = 0.0;
  return l;
}
static inline pt_t line_direction(const line_t *l) { return pt_unit(pt_sub(l->en, l->st)); } /* :239 */
static inline double line_distance_to_point(const line_t *l, pt_t p) {            /* :244 */
  double d = l->a * p.x + l->b * p.y + l->c;
  double dn = sqrt(l->a * l->a + l->b * l->b);
  return d / dn;
}
static inline pt_t line_normal(const line_t *l) { return pt_unit(P(-l->a, -l->b)); } /* :251 */

/* get_relative_location_points, :288 */
static int line_rel_loc(const line_t *l, pt_t p1, pt_t p2) {
  pt_t d = pt_unit(pt_sub(p2, p1));
  pt_t ld = line_direction(l);
  double c = pt_cosine(d, ld);
  if (c > 1 - TOL && c < 1 + TOL) return 1;
  if (c > -1 - TOL && c < -1 + TOL) return -1;
  return 0;
}
/* is_on_segment, :310 */
static int line_on_segment(const line_t *l, pt_t p) {
  double d1 = pt_distance(l->st, p);
  double d2 = pt_distance(l->en, p);
  double d3 = pt_distance(l->st, l->en);
  double ds = d1 + d2;
  return (ds <= d3 + TOL) && (ds >= d3 - TOL);
}
/* get_intersection, :327; returns 0 for None */
static int line_intersection(const line_t *s, const line_t *l2, pt_t *out) {
  double nr = l2->a * s->c - s->a * l2->c;
  double dr = s->a * l2->b - l2->a * s->b;
  if (dr == 0) return 0;
  double y = nr / dr, x;
  if (s->a == 0) x = -(l2->c + l2->b * y) / l2->a;
  else x = -(s->c + s->b * y) / s->a;
  *out = P(x, y);
  return 1;
}
/* get_intersection_ray, :347 */
static int line_intersection_ray(const line_t *s, const line_t *ray, pt_t *out) {
  pt_t p;
  if (!line_intersection(s, ray, &p)) return 0;
  if (line_rel_loc(ray, ray->st, p) != 1) return 0;
  *out = p;
  return 1;
}
/* get_normal_towards_point, :258 */
static pt_t line_normal_towards_point(const line_t *l, pt_t p) {
  pt_t n = line_normal(l);
  line_t ray = line_make(p, pt_add(p, n));
  pt_t ip;
  if (!line_intersection_ray(l, &ray, &ip)) return n;
  return pt_scale(n, -1);
}

/* ---- Circle.intersect_moving_circle, geometry.py:400-466 ----------------- */
/* returns 1 with (tcol, nrml) when a contact exists, 0 for (inf,None,None);
 * *err != 0 when the reference would have raised */
static int circle_intersect_moving(double r1, pt_t p1, double r2, pt_t p2, pt_t v21,
                                   double *tcol, pt_t *nrml, int *err) {
  pt_t p11 = pt_sub(p1, p1);
  pt_t p21 = pt_sub(p2, p1);
  double R = r1 + r2;
  double dBall = pt_distance(p21, p11);
  if (dBall >= (R - 0.1) && dBall <= R) dBall = (R - dBall) + dBall + 1e-6;       /* :411 */
  if (!(dBall >= R)) { *err = ST_ASSERT_OVERLAP; return 0; }                      /* :413 */
  line_t lv = line_make(p21, pt_add(p21, v21));
  double dl = fabs(line_distance_to_point(&lv, p11));                              /* :393 */
  if (!(dl <= R)) return 0;                                                       /* :417 */
  pt_t c1c2 = pt_scale(pt_sub(p21, p11), -1);                                     /* :425 */
  double theta = pt_angle_between(c1c2, v21, err);                                /* :435 */
  if (*err) return 0;
  if (!(fabs(theta) < M_PI / 2)) { *err = ST_ASSERT_THETA; return 0; }            /* :436-438 */
  theta = fabs(theta);
  double dist;
  if (theta == 0) dist = dBall - R;                                               /* :440 */
  else {
    double sinC = dBall / (R / o_sin(theta));                                     /* :443 */
    if (!(sinC >= -1.0 && sinC <= 1.0)) { *err = ST_MATH_DOMAIN; return 0; }
    double thetaC = M_PI - o_asin(sinC);
    double thetaA = M_PI - (thetaC + theta);
    if (!(thetaA >= 0)) { *err = ST_ASSERT_THETA_A; return 0; }                   /* :447 */
    dist = (R / o_sin(theta)) * o_sin(thetaA);                                    /* :448 */
  }
  double speed = pt_mag(v21);
  if (speed == 0) return 0;                                                       /* :453 */
  pt_t vDir = pt_unit(v21);
  *tcol = dist / speed;
  pt_t newP2 = pt_add(p21, pt_mul(vDir, dist));                                   /* :460 */
  line_t lc = line_make(newP2, p11);                                              /* :463 */
  *nrml = line_normal(&lc);
  return 1;
}

/* ---- the world ------------------------------------------------------------ */
typedef struct {
  int n_balls, n_walls;
  const int32_t *order;   /* n_walls + n_balls entries: q < n_walls -> wall q, else ball q - n_walls */
  double *pos, *vel;  /* [n_balls][2] */
  const double *rad, *mass; /* [n_balls]; rad < 0 marks an empty slot */
  const double *wall; /* [n_walls][4 vertices][2] */
  double *tcol;       /* [n_balls] */
  double *fvel;       /* [n_balls][2]  futureVel_ */
  double *nrml;       /* [n_balls][2]  nrmlCol_ */
  int32_t *partner;   /* [n_balls]  object index as in `order`, -1 for (None, None) */
  int32_t *pending;   /* colQueue_ non-empty: every ball is rescanned at the next loop head */
  int32_t *status, *status_step;
  double dt, gx, gy;
  int max_substeps;
} world_t;

/* dynamics.get_toc_ball_wall, dynamics.py:8-63 */
static double toc_ball_wall(pt_t pos, pt_t vel, double r, const double *wv, pt_t *nrmlCol, int *err) {
  double tCol = INFINITY;
  for (int e = 0; e < 4; ++e) {
    pt_t v0 = P(wv[2 * e], wv[2 * e + 1]);
    pt_t v1 = P(wv[2 * ((e + 1) & 3)], wv[2 * ((e + 1) & 3) + 1]);
    line_t l = line_make(v0, v1);                                                 /* geometry.py:499-502 */
    pt_t nrml = line_normal_towards_point(&l, pos);                               /* :22 */
    nrml = pt_scale(nrml, -1);                                                    /* :23 */
    double speed = pt_dot(vel, nrml);                                             /* :24 */
    if (speed <= 0) continue;
    pt_t velOrth = pt_sub(vel, pt_mul(nrml, speed));                              /* :28 */
    line_t ray = line_make(pos, pt_add(pos, nrml));                               /* :31 */
    pt_t ip;
    if (!line_intersection_ray(&l, &ray, &ip)) { *err = ST_ASSERT_INTPOINT; return tCol; } /* :34 */
    double distCenter = pt_distance(pos, ip);
    double dist = distCenter - r;
    if (dist < 0) {                                                               /* :38 */
      if (line_on_segment(&l, ip)) { *err = ST_TRAP_AMISS; return tCol; }         /* :43-45 */
      continue;
    }
    double t = dist / speed;                                                      /* :48 */
    pt_t lp = pt_add(ip, pt_mul(velOrth, t));                                     /* :51 */
    if (!line_on_segment(&l, lp)) continue;
    if (t < tCol) { tCol = t; *nrmlCol = pt_scale(nrml, -1); }                    /* :55-58 */
  }
  return tCol;
}

/* dynamics.get_toc_ball_ball, dynamics.py:67-136; writes futureVel of ball 1 */
static double toc_ball_ball(world_t *w, int b1, int b2, pt_t *nrmlCol, int *err) {
  pt_t pos1 = P(w->pos[2 * b1], w->pos[2 * b1 + 1]), vel1 = P(w->vel[2 * b1], w->vel[2 * b1 + 1]);
  pt_t pos2 = P(w->pos[2 * b2], w->pos[2 * b2 + 1]), vel2 = P(w->vel[2 * b2], w->vel[2 * b2 + 1]);
  pt_t relVel = pt_sub(vel2, vel1);
  pt_t colDir = pt_unit(pt_sub(pos2, pos1));
  double speed = -pt_dot(relVel, colDir);                                         /* :84 */
  if (speed <= 0) return INFINITY;
  double tCol; pt_t n;
  if (!circle_intersect_moving(w->rad[b1], pos1, w->rad[b2], pos2, relVel, &tCol, &n, err))
    return INFINITY;
  pt_t vOth1 = pt_project(vel1, n), vOth2 = pt_project(vel2, n);                  /* :97-98 */
  pt_t vCol1 = pt_sub(vel1, vOth1), vCol2 = pt_sub(vel2, vOth2);
  double m1 = w->mass[b1], m2 = w->mass[b2];
  pt_t v1n = pt_mul(pt_add(pt_mul(vCol1, m1 - m2), pt_mul(vCol2, 2 * m2)), 1.0 / (m1 + m2)); /* :125 */
  pt_t vel1New = pt_add(v1n, vOth1);                                              /* :127 */
  w->fvel[2 * b1] = vel1New.x; w->fvel[2 * b1 + 1] = vel1New.y;                   /* :129 */
  *nrmlCol = n;
  return tCol;
}

/* Dynamics.time_to_collide_all, primitives.py:720-733 */
static void time_to_collide_all(world_t *w, int b, int *err) {
  w->tcol[b] = INFINITY;
  w->partner[b] = -1;
  pt_t pos = P(w->pos[2 * b], w->pos[2 * b + 1]), vel = P(w->vel[2 * b], w->vel[2 * b + 1]);
  int nobj = w->n_walls + w->n_balls;
  for (int k = 0; k < nobj; ++k) {
    int q = w->order[k];
    double toc; pt_t n = P(0, 0);
    if (q < w->n_walls) {
      toc = toc_ball_wall(pos, vel, w->rad[b], w->wall + 8 * q, &n, err);
    } else {
      int b2 = q - w->n_walls;
      if (b2 == b || w->rad[b2] < 0) continue;
      toc = toc_ball_ball(w, b, b2, &n, err);
    }
    if (*err) return;
    if (toc < w->tcol[b]) {                                                       /* :729 */
      w->tcol[b] = toc; w->partner[b] = q;
      w->nrml[2 * b] = n.x; w->nrml[2 * b + 1] = n.y;
    }
  }
}

/* Dynamics.move_object, primitives.py:600-611 */
static void move_object(world_t *w, int b, double t, int *err) {
  pt_t pos = P(w->pos[2 * b], w->pos[2 * b + 1]), vel = P(w->vel[2 * b], w->vel[2 * b + 1]);
  pt_t g = P(w->gx, w->gy);
  pos = pt_add(pt_add(pos, pt_mul(vel, t)), pt_mul(g, 0.5 * t * t));              /* :604 */
  vel = pt_add(vel, pt_mul(g, t));                                                /* :606 */
  w->pos[2 * b] = pos.x; w->pos[2 * b + 1] = pos.y;
  w->vel[2 * b] = vel.x; w->vel[2 * b + 1] = vel.y;
  w->tcol[b] = w->tcol[b] - t;
  if (!(w->tcol[b] > -1e-8)) *err = ST_ASSERT_TCOL_NEG;                           /* :610 */
}

typedef struct { int32_t world, step, ball, partner; } event_t;

typedef struct {
  event_t *buf; int64_t cap; int64_t n;
} evlog_t;

/* Dynamics.step, primitives.py:628-656 (+ step_object :614, resolve_collision :659) */
static double g_friction = 0.0; /* engine extension (ppb_config.friction); 0 = the reference */

static void world_step(world_t *w, int world_id, int step_idx, evlog_t *ev) {
  if (*w->status != ST_OK) return;
  int err = 0;
  if (g_friction != 0.0) {
    /* NOT in the reference: semi-implicit Euler friction as the engine defines it.  At the
     * start of a step every velocity -- the cached futureVel_ included -- is scaled by s = max(0, 1 - mu*dt);
     * a uniform scaling leaves every contact geometry unchanged and stretches every cached tCol_ by 1/s. */
    double s = fmax(0.0, 1.0 - g_friction * w->dt);
    for (int b = 0; b < w->n_balls; ++b) {
      if (w->rad[b] < 0) continue;
      w->vel[2 * b] = w->vel[2 * b] * s; w->vel[2 * b + 1] = w->vel[2 * b + 1] * s;
      w->fvel[2 * b] = w->fvel[2 * b] * s; w->fvel[2 * b + 1] = w->fvel[2 * b + 1] * s;
      w->tcol[b] = s > 0 ? w->tcol[b] / s : INFINITY;
    }
  }
  double tStep = 0;
  int iters = 0;
  for (;;) {
    if (*w->pending) {                                                            /* :634-638 */
      *w->pending = 0;
      for (int b = 0; b < w->n_balls && !err; ++b)
        if (w->rad[b] >= 0) time_to_collide_all(w, b, &err);
      if (err) break;
    }
    double mnt = INFINITY;                                                        /* :641-646 */
    for (int b = 0; b < w->n_balls; ++b)
      if (w->rad[b] >= 0 && w->tcol[b] < mnt) mnt = w->tcol[b];
    double rem = w->dt - tStep;
    double t = (rem < mnt) ? rem : mnt;                                           /* :647 */
    if (t <= 0) break;
    if (++iters > w->max_substeps) { err = ST_ITER_LIMIT; break; }
    for (int b = 0; b < w->n_balls && !err; ++b) {                                /* :653 */
      if (w->rad[b] < 0) continue;
      if (w->tcol[b] <= t) {                                                      /* :617 */
        if (!(w->tcol[b] == t)) { err = ST_ASSERT_TCOL_EQ; break; }               /* :664 */
        if (ev->n < ev->cap) {
          event_t e = {world_id, step_idx, b, w->partner[b]};
          ev->buf[ev->n] = e;
        }
        ev->n++;
        move_object(w, b, w->tcol[b], &err);                                      /* :665 */
        if (err) break;
        int q = w->partner[b];
        if (q >= 0 && q < w->n_walls) {                                           /* :666-674 */
          pt_t n = P(w->nrml[2 * b], w->nrml[2 * b + 1]);
          pt_t v = pt_reflect_normal(n, P(w->vel[2 * b], w->vel[2 * b + 1]));
          w->vel[2 * b] = v.x; w->vel[2 * b + 1] = v.y;
        } else {                                                                  /* :677-681 */
          w->vel[2 * b] = w->fvel[2 * b]; w->vel[2 * b + 1] = w->fvel[2 * b + 1];
        }
        *w->pending = 1;                                                          /* :697-700 */
      } else {
        move_object(w, b, t, &err);                                               /* :621 */
      }
    }
    if (err) break;
    tStep += t;                                                                   /* :655 */
  }
  if (err) { *w->status = err; *w->status_step = step_idx; }
}

/* ---- batch entry points (ctypes) ------------------------------------------- */
/*
 * Arrays are world-major C-contiguous.  `traj` (may be NULL) receives the
 * positions after each step: [n_steps][n_worlds][n_balls][2], i.e. what
 * DataSaver.save_sequence records (dataio.py:123-128).  Events are one record
 * per resolve_collision call in call order within a world (worlds are processed
 * in index order when n_threads == 1).  Returns the number of events produced
 * (which may exceed ev_cap; only the first ev_cap are stored).
 */
int64_t oracle_step_batch(int n_worlds, int n_balls, int n_walls, const int32_t *order,
                          double *pos, double *vel, const double *rad, const double *mass,
                          const double *wall, double *tcol, double *fvel, double *nrml,
                          int32_t *partner, int32_t *pending, int32_t *status, int32_t *status_step,
                          double dt, double gx, double gy, int step0, int n_steps,
                          double *traj, int32_t *ev_buf, int64_t ev_cap, int trig_mode, int n_threads,
                          int max_substeps) {
  g_trig_mode = trig_mode;
  int64_t n_events = 0;
  /* the event list is ordered (world-major, call order inside a world), so it is
   * only materialised single-threaded; with n_threads > 1 events are counted */
  if (ev_buf) n_threads = 1;
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads) reduction(+ : n_events)
  for (int wi = 0; wi < n_worlds; ++wi) {
    evlog_t ev = {ev_buf ? (event_t *)ev_buf + n_events : NULL, ev_buf ? ev_cap - n_events : 0, 0};
    if (ev.cap < 0) ev.cap = 0;
    world_t w;
    w.n_balls = n_balls; w.n_walls = n_walls; w.order = order;
    w.pos = pos + (size_t)wi * n_balls * 2; w.vel = vel + (size_t)wi * n_balls * 2;
    w.rad = rad + (size_t)wi * n_balls; w.mass = mass + (size_t)wi * n_balls;
    w.wall = wall + (size_t)wi * n_walls * 8;
    w.tcol = tcol + (size_t)wi * n_balls; w.fvel = fvel + (size_t)wi * n_balls * 2;
    w.nrml = nrml + (size_t)wi * n_balls * 2; w.partner = partner + (size_t)wi * n_balls;
    w.pending = pending + wi; w.status = status + wi; w.status_step = status_step + wi;
    w.dt = dt; w.gx = gx; w.gy = gy;
    w.max_substeps = max_substeps > 0 ? max_substeps : (1 << 20);
    for (int s = 0; s < n_steps; ++s) {
      world_step(&w, wi, step0 + s, &ev);
      if (traj) memcpy(traj + ((size_t)s * n_worlds + wi) * n_balls * 2, w.pos,
                       sizeof(double) * n_balls * 2);
    }
    n_events += ev.n;
  }
  return n_events;
}

void oracle_set_friction(double mu) { g_friction = mu; }

/* single geometry helpers exported for the known-answer tests (unit_tests.py) */
int oracle_line_intersection(const double *l1, const double *l2, double *out) {
  line_t a = line_make(P(l1[0], l1[1]), P(l1[2], l1[3]));
  line_t b = line_make(P(l2[0], l2[1]), P(l2[2], l2[3]));
  pt_t p;
  if (!line_intersection(&a, &b, &p)) return 0;
  out[0] = p.x; out[1] = p.y;
  return 1;
}
int oracle_line_intersection_ray(const double *l1, const double *ray, double *out) {
  line_t a = line_make(P(l1[0], l1[1]), P(l1[2], l1[3]));
  line_t b = line_make(P(ray[0], ray[1]), P(ray[2], ray[3]));
  pt_t p;
  if (!line_intersection_ray(&a, &b, &p)) return 0;
  out[0] = p.x; out[1] = p.y;
  return 1;
}
void oracle_reflect_normal(const double *nrml, const double *v, double *out) {
  pt_t r = pt_reflect_normal(P(nrml[0], nrml[1]), P(v[0], v[1]));
  out[0] = r.x; out[1] = r.y;
}
void oracle_normal_towards_point(const double *l1, const double *p, double *out) {
  line_t a = line_make(P(l1[0], l1[1]), P(l1[2], l1[3]));
  pt_t n = line_normal_towards_point(&a, P(p[0], p[1]));
  out[0] = n.x; out[1] = n.y;
}
void oracle_line_canonical(const double *l1, double *abc) {
  line_t a = line_make(P(l1[0], l1[1]), P(l1[2], l1[3]));
  abc[0] = a.a; abc[1] = a.b; abc[2] = a.c;
}
double oracle_trig(int which, double x, int mode) {
  g_trig_mode = mode;
  return which == 0 ? o_sin(x) : which == 1 ? o_asin(x) : o_acos(x);
}
