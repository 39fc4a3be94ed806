// ppb_kernels.cuh -- the per-step kernels of the batched billiards engine (sm_100a).
//
//   advance_kernel      : n x Dynamics.step (primitives.py:600-733 + dynamics.py:8-136) for every world of the
//                         batch in ONE launch.  A warp (or a team of warps) keeps a world in registers for all n
//                         steps: event-free steps are Dynamics.move_object and nothing else, the others run the
//                         event-driven loop with the world's balls and wall edges staged in shared memory, the
//                         (ball x object) TOC tests spread over the lanes and reduced with warp shuffles
//                         (lexicographic (t, insertion rank) arg-min).
//   render_tile_kernel  : World.generate_image (primitives.py:991-1008) -- in ppb_render.cuh.
#pragma once
#include "ppb_math.cuh"

namespace ppb {

// Programmatic dependent launch (sm_90+): a kernel launched with programmatic stream serialisation may start
// while its predecessor in the stream drains; it must not touch global memory before pdl_wait().  Both are
// no-ops for a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = (u < v) ? u : v;
  }
  return v;
}

// Dynamics.move_object, primitives.py:604-606
template <typename T>
__device__ __forceinline__ void move_ball(V2<T> &pos, V2<T> &vel, T t, V2<T> g) {
  T k = (T(0.5) * t) * t;
  pos = mk<T>((pos.x + vel.x * t) + g.x * k, (pos.y + vel.y * t) + g.y * k);
  vel = mk<T>(vel.x + g.x * t, vel.y + g.y * t);
}

// ------------------------------------------------------------------------------------
// K2: collide
//
// One warp advances one world.  The 32 lanes are LB x G: LB = next power of two >= n_balls
// ball columns, G = 32 / LB groups.  Ball state is replicated in the registers of the G lanes
// of its column; the time-of-collision scan of a ball (Dynamics.time_to_collide_all,
// primitives.py:720-733) is spread over its G lanes -- lane (ball i, group g) tests the scan
// items g, g + G, g + 2G, ... (an item = one wall edge or one other ball, numbered in the
// reference's visiting order) -- and merged with xor-shuffles:
//   * collision time: lexicographic (t, item number) minimum == the reference's strict '<'
//     running minimum (first object / first edge in order wins ties);
//   * futureVel_: the ball-ball item with the largest number that got as far as writing it
//     (dynamics.py:129: every candidate partner overwrites it);
//   * a raised assertion: the item with the smallest number.
// All step-loop control flow is warp-uniform (one world per warp).  Worlds are handed out
// through an atomic counter so that a world with a long chain of sub-steps delays only itself.
// ------------------------------------------------------------------------------------
template <typename T>
struct WarpSmem {
  typename Vec2<T>::type pos[32];
  typename Vec2<T>::type vel[32];
  typename Vec2<T>::type rm[32];
  typename MathOf<T>::type::Edge edge[PPB_MAX_WALLS * 4];
};

#define PPB_COLLIDE_WARPS 4
#define PPB_K_NONE 0x7fff

template <typename T, int LW>
__device__ __forceinline__ T seg_min(T v) {
#pragma unroll
  for (int o = LW >> 1; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = (u < v) ? u : v;
  }
  return v;
}

// minimum over the warp.  fp32: one redux.sync on an order-preserving integer image of the floats
// (sm_80+); fp64: xor-shuffle tree.  Inputs are never NaN.
__device__ __forceinline__ float warp_min_all(float v) {
  const uint32_t b = __float_as_uint(v);
  const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  const uint32_t m = __reduce_min_sync(0xffffffffu, key);
  return __uint_as_float((m & 0x80000000u) ? (m & 0x7fffffffu) : ~m);
}
__device__ __forceinline__ double warp_min_all(double v) { return seg_min<double, 32>(v); }

// status of the first ball (in order) that raised: every lane of a ball column holds the same err,
// lanes 0..LB-1 are the balls in order
template <int LB>
__device__ __forceinline__ int first_error(int err) {
  const unsigned em = __ballot_sync(0xffffffffu, err != 0) & (LB == 32 ? 0xffffffffu : ((1u << LB) - 1u));
  if (em == 0u) return 0;
  return __shfl_sync(0xffffffffu, err, __ffs(em) - 1);
}

__device__ __forceinline__ float shfl_xor_t(float v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double shfl_xor_t(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// partial scan result of one ball, exchanged between the warps of a team through shared memory
template <typename T>
struct ScanPart {
  T t, nx, ny, fx, fy;
  int k, partner, kfv, kerr, err;
};
template <typename T, int TW>
struct TeamSmem {
  WarpSmem<T> w;                                  // ball / edge staging shared by the team
  ScanPart<T> part[TW > 1 ? TW : 1][TW > 1 ? 32 : 1];
  int ticket;
};

template <int TW>
__device__ __forceinline__ void team_sync(int team) {
  if (TW > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(TW * 32) : "memory");
  else __syncwarp();
}

// TW warps work on one world (a "team"): TW = 1 for up to 8 balls, 2 for 16, 4 for 32, so that a lane
// never has more than a handful of scan items.  Every warp of a team holds the whole (replicated) ball
// state and takes the same decisions; only the scan is split, and its partial results are exchanged
// through shared memory.
//
// The team keeps its world for all P.n_steps steps of the launch: positions, velocities, tCol_ and the
// world's min tCol_ live in registers from step to step.  A step whose header says "no queued rescan and
// min tCol_ > dt" is Dynamics.move_object for every ball (primitives.py:600-611) and nothing else; any
// other step runs the event loop below.  The collision caches the event loop needs (nrmlCol_, futureVel_,
// objCol_, the wall edges) are fetched when the first such step comes up.  HBM sees one read and one
// write of the state per launch, plus the trajectory rows / position snapshots the caller asked for.
template <typename T, int LB, int TW>
__global__ void __launch_bounds__(PPB_COLLIDE_WARPS * 32)
advance_kernel(StatePtrs S, Layout L, StepParams P, T *traj) {
  typedef typename MathOf<T>::type M;
  typedef typename Vec2<T>::type T2;
  typedef typename Vec4<T>::type T4;
  constexpr int TEAMS = PPB_COLLIDE_WARPS / TW;
  __shared__ TeamSmem<T, TW> smem[TEAMS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int team = warp / TW, wt = warp % TW;  // team in the CTA, warp in the team
  TeamSmem<T, TW> &ts = smem[team];
  WarpSmem<T> &sm = ts.w;
  const unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / LB;                   // lanes per ball inside a warp
  constexpr int GT = G * TW;                   // scan groups per ball in the team
  const int bi = lane & (LB - 1), grp = lane / LB;
  const int gid = wt * G + grp;                // this lane's scan group
  const bool writer = (wt == 0) && (grp == 0); // the lane that owns ball bi's global state
  const int nB = L.n_balls, nWl = L.n_walls;
  pdl_launch_dependents();   // the next launch of the stream may take its places; it waits for this grid to finish
  pdl_wait();                // the previous launch (state, ticket counter) is complete and visible
  const uint32_t k0 = (uint32_t)P.step_abs;
  const T dt = (T)P.dt;
  const V2<T> g = mk<T>((T)P.gx, (T)P.gy);
  const bool has_g = (P.gx != 0.0) || (P.gy != 0.0);
  const T INF = Lim<T>::inf();
  // friction extension (0 = the reference): every velocity is scaled by s = 1 - friction*dt at the start of
  // the step (semi-implicit Euler: the step then moves with the new velocity).  A uniform scaling of all
  // velocities leaves every contact geometry unchanged and stretches every cached time of collision by
  // 1/s (the cached futureVel_ shrinks by s), so the event-driven caches stay exact.
  const bool fric = (P.friction != 0.0) && !P.scan_only;
  const T fs = fric ? (T)fmax(0.0, 1.0 - P.friction * P.dt) : T(1);

  if (blockIdx.x == 0 && threadIdx.x == 0) S.list_count[2 + ((P.parity ^ 1) & 1)] = 0;   // next launch's tickets
  int *ticket = S.list_count + 2 + (P.parity & 1);
  const int count = S.n_worlds;
  const long long total = (long long)count * nB;

  Hdr<T> *hdr = reinterpret_cast<Hdr<T> *>(S.hdr);
  T2 *gpos = reinterpret_cast<T2 *>(S.pos);
  T2 *gvel = reinterpret_cast<T2 *>(S.vel);
  T2 *grm = reinterpret_cast<T2 *>(S.rm);
  T4 *gaux = reinterpret_cast<T4 *>(S.aux);
  T *gtcol = reinterpret_cast<T *>(S.tcol);
  const T2 *gwall = reinterpret_cast<const T2 *>(S.wall);

  unsigned long long cnt_iters = 0, cnt_rescans = 0, cnt_slow = 0;
  int idx = 0;
  if (TW == 1) {
    if (lane == 0) idx = atomicAdd(ticket, 1);
    idx = __shfl_sync(FULL, idx, 0);
  } else {
    if (wt == 0 && lane == 0) ts.ticket = atomicAdd(ticket, 1);
    team_sync<TW>(team);
    idx = ts.ticket;
    team_sync<TW>(team);
  }

  // a ticket is P.ticket_worlds consecutive worlds
  const int tkw = P.ticket_worlds;
  while ((long long)idx * tkw < count) {
    int idx_next = 0;
    if (wt == 0 && lane == 0) idx_next = atomicAdd(ticket, 1);   // its latency hides behind these worlds' work
    const int w_first = idx * tkw;
    const int w_end = min(count, w_first + tkw);
    for (int w = w_first; w < w_end; ++w) {
    const bool has_ball = bi < nB;
    const long long base = (long long)w * nB;

    // ---- the world's moving state: replicated over the lanes of a ball column ----
    const Hdr<T> h = hdr[w];
    V2<T> pos = mk<T>(0, 0), vel = pos, nrm = pos, fv = pos;
    T tc = INF, rad = T(-1), mass = T(1);
    int partner = -1;
    if (has_ball) {
      const T2 p = gpos[base + bi], v = gvel[base + bi], q = grm[base + bi];
      pos = mk<T>(p.x, p.y);
      vel = mk<T>(v.x, v.y);
      rad = q.x;
      mass = q.y;
      tc = gtcol[base + bi];
    }
    const bool active = has_ball && rad >= T(0);
    int status = (int)PPB_STATUS_OF(h.flags);
    bool pending = (h.flags & PPB_FLAG_PENDING) != 0u;
    T mnt = h.tmin;
    bool warm = false, staged = false, rescanned = false, vel_dirty = has_g || fric;
    uint32_t nev_base = 0u, nev_local = 0u;
    int s_done = 0;
    const bool skip = P.scan_only && (!pending || status != 0);   // ppb_scan_async only drains non-empty queues
    // where this lane's ball goes in the trajectory rows / snapshot slots of the steps to come
    const bool rec_traj = traj != nullptr && has_ball && writer && !P.scan_only;
    T2 *traj_p = reinterpret_cast<T2 *>(traj) + (long long)P.step_off * total + base + bi;
    T2 *snap_p = reinterpret_cast<T2 *>(S.snap) + base + bi;
    int next_draw = (S.snap && P.draw_every > 0 && !P.scan_only) ? P.draw_first : 0x7fffffff;

    for (int s = 0; s < P.n_steps && !skip; ++s) {
      if (status == 0) {
        const T mnt_eff = fric ? (fs > T(0) ? mnt / fs : INF) : mnt;
        if (!pending && !P.scan_only && mnt_eff > dt) {
          // ---- event-free step: Dynamics.move_object for every ball (primitives.py:600-611) ----
          if (fric) {
            if (!warm) {   // the cached futureVel_ shrinks with the velocities
              if (has_ball) {
                const T4 a = ld4(gaux + base + bi);
                nrm = mk<T>(a.x, a.y);
                fv = mk<T>(a.z, a.w);
                partner = S.partner[base + bi];
              }
              warm = true;
            }
            vel = vel * fs;
            fv = fv * fs;
            tc = fs > T(0) ? tc / fs : INF;
          }
          if (active) {
            move_ball(pos, vel, dt, g);
            tc = tc - dt;
          }
          mnt = mnt_eff - dt;   // min_b (tCol_b - dt) == (min_b tCol_b) - dt: rounding is monotone
          ++s_done;
        } else {
          // ---- the event-driven loop of Dynamics.step (primitives.py:628-700) ----
          if (!warm) {
            if (has_ball) {
              const T4 a = ld4(gaux + base + bi);
              nrm = mk<T>(a.x, a.y);
              fv = mk<T>(a.z, a.w);
              partner = S.partner[base + bi];
            }
            warm = true;
          }
          if (!staged) {
            // first event step of this world in this launch: fetch the event sequence number, stage the wall
            // edges and radii / masses (the staging area is shared by the team: its previous readers must be done)
            nev_base = S.nev[w];
            team_sync<TW>(team);
            for (int e = wt * 32 + lane; e < nWl * 4; e += TW * 32) {
              const T2 *wp = gwall + ((long long)w * nWl + (e >> 2)) * 4;
              const T2 a = wp[e & 3], b = wp[(e + 1) & 3];
              sm.edge[e] = M::make_edge(mk<T>(a.x, a.y), mk<T>(b.x, b.y));
            }
            if (writer) {
              T2 q;
              q.x = rad; q.y = mass;
              sm.rm[bi] = q;
            }
            staged = true;
          }
          if (fric) {
            vel = vel * fs;
            fv = fv * fs;
            tc = fs > T(0) ? tc / fs : INF;
          }
          vel_dirty = true;
          ++cnt_slow;
          const uint32_t k = k0 + (uint32_t)s;
          bool done = false;
          int err = 0;
          T tStep = T(0);
          int iters = 0;
          for (;;) {
            if (pending) {
              // colQueue_ drain: time_to_collide_all for every ball of the world (primitives.py:634-638)
              team_sync<TW>(team);
              if (writer) {
                T2 p, v;
                p.x = pos.x; p.y = pos.y; v.x = vel.x; v.y = vel.y;
                sm.pos[bi] = p;
                sm.vel[bi] = v;
              }
              team_sync<TW>(team);
              tc = INF;
              partner = -1;
              int kbest = PPB_K_NONE, kfv = -1, kerr = PPB_K_NONE;
              V2<T> nb = nrm, fb = fv;
              if (active) {
                // The lane's share of the scan items, edges first, then balls.  The order of evaluation is
                // free: every candidate carries its position k in the reference's visiting order and ties
                // are broken on k, so the result equals the sequential strict-'<' scan.  Type-pure loops
                // with independent trips let the tests of several items overlap in the pipeline.
                const int nE = nWl << 2;
#pragma unroll 4
                for (int e = gid; e < nE; e += GT) {
                  // one edge of dynamics.get_toc_ball_wall (dynamics.py:20-59)
                  const int kk = L.edge_k[e];
                  T te;
                  V2<T> ne;
                  const int ee = M::edge_toc(pos, vel, rad, sm.edge[e], te, ne);
                  if (ee) {
                    if (kk < kerr) { kerr = kk; err = ee; }
                  } else if (te < tc || (te == tc && kk < kbest && kbest != PPB_K_NONE)) {
                    tc = te; kbest = kk; partner = e >> 2; nb = ne;
                  }
                }
#pragma unroll 2
                for (int b2 = gid; b2 < nB; b2 += GT) {
                  const T2 q2 = sm.rm[b2];
                  if (b2 == bi || q2.x < T(0)) continue;
                  const int kk = L.ball_k[b2];
                  const T2 p2 = sm.pos[b2], v2 = sm.vel[b2];
                  const Toc<T> r = M::ball(pos, vel, rad, mass, mk<T>(p2.x, p2.y), mk<T>(v2.x, v2.y), q2.x, q2.y);
                  if (r.err) {
                    if (kk < kerr) { kerr = kk; err = r.err; }
                    continue;
                  }
                  if (r.has_fv && kk > kfv) { fb = r.fv; kfv = kk; }   // futureVel_: last partner in order (dynamics.py:129)
                  if (r.t < tc || (r.t == tc && kk < kbest && kbest != PPB_K_NONE)) {
                    tc = r.t; kbest = kk; partner = nWl + b2; nb = r.n;
                  }
                }
              }
              // merge the partial scans of every ball: inside the warp ...
#pragma unroll
              for (int o = LB; o < 32; o <<= 1) {
                const T ot = shfl_xor_t(tc, o);
                const int ok = __shfl_xor_sync(FULL, kbest, o), op = __shfl_xor_sync(FULL, partner, o);
                const T onx = shfl_xor_t(nb.x, o), ony = shfl_xor_t(nb.y, o);
                if (ot < tc || (ot == tc && ok < kbest)) {
                  tc = ot; kbest = ok; partner = op; nb = mk<T>(onx, ony);
                }
                const int okf = __shfl_xor_sync(FULL, kfv, o);
                const T ofx = shfl_xor_t(fb.x, o), ofy = shfl_xor_t(fb.y, o);
                if (okf > kfv) { kfv = okf; fb = mk<T>(ofx, ofy); }
                const int oke = __shfl_xor_sync(FULL, kerr, o), oe = __shfl_xor_sync(FULL, err, o);
                if (oke < kerr) { kerr = oke; err = oe; }
              }
              // ... and across the warps of the team
              if (TW > 1) {
                if (grp == 0) {
                  ScanPart<T> sp;
                  sp.t = tc; sp.nx = nb.x; sp.ny = nb.y; sp.fx = fb.x; sp.fy = fb.y;
                  sp.k = kbest; sp.partner = partner; sp.kfv = kfv; sp.kerr = kerr; sp.err = err;
                  ts.part[wt][bi] = sp;
                }
                team_sync<TW>(team);
#pragma unroll
                for (int u = 0; u < TW; ++u) {
                  if (u == wt) continue;
                  const ScanPart<T> o = ts.part[u][bi];
                  if (o.t < tc || (o.t == tc && o.k < kbest)) {
                    tc = o.t; kbest = o.k; partner = o.partner; nb = mk<T>(o.nx, o.ny);
                  }
                  if (o.kfv > kfv) { kfv = o.kfv; fb = mk<T>(o.fx, o.fy); }
                  if (o.kerr < kerr) { kerr = o.kerr; err = o.err; }
                }
              }
              // an item after a raising one is never visited by the reference; the world halts anyway
              if (kbest != PPB_K_NONE) nrm = nb;
              if (kfv >= 0) fv = fb;
              rescanned = true;
              pending = false;
              ++cnt_rescans;
              // the first ball (in order) whose scan raised decides the status
              const int e = first_error<LB>(err);
              if (e) {
                status = e;
                done = true;
              }
            }
            if (P.scan_only) {   // ppb_scan_async: the caches are up to date, nothing moves
              if (!done) mnt = warp_min_all(active ? tc : INF);
              break;
            }
            // mntCol / t  (primitives.py:641-650)
            const T m = warp_min_all(active ? tc : INF);   // every group holds the same ball columns
            T t = T(0);
            if (!done) {
              mnt = m;
              const T rem = dt - tStep;
              t = (rem < mnt) ? rem : mnt;
              if (t <= T(0)) done = true;
              else if (++iters > P.max_substeps) {
                status = PPB_ST_ITER_LIMIT;
                done = true;
              }
            }
            if (done) break;
            const bool hit = active && (tc <= t);   // step_object, primitives.py:617
            const unsigned evm = __ballot_sync(FULL, hit && grp == 0);
            if (evm) {
              // one record per resolve_collision call, in ball order (written by the first warp of the team)
              const int n_ev = __popc(evm);
              if (S.event_capacity > 0 && wt == 0) {
                unsigned long long slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(S.counters + 0, (unsigned long long)n_ev);
                slot0 = __shfl_sync(FULL, slot0, 0);
                if (hit && grp == 0) {
                  const int ord = __popc(evm & ((1u << lane) - 1u));
                  const unsigned long long slot = slot0 + ord;
                  if ((long long)slot < S.event_capacity) {
                    ppb_event ev;
                    ev.world = w;
                    ev.seq = (int32_t)(nev_base + nev_local + ord);
                    ev.step = (int32_t)k;
                    ev.ball = (int16_t)bi;
                    ev.partner = (int16_t)partner;
                    S.events[slot] = ev;
                  }
                }
              }
              nev_local += n_ev;
              pending = true;   // add_all_dynamic_collision_queue, primitives.py:697-700
            }
            if (active) {
              if (hit) {  // resolve_collision, primitives.py:659-681
                if (!(tc == t)) err = PPB_ST_ASSERT_TCOL_EQ;
                move_ball(pos, vel, tc, g);
                tc = tc - tc;
                if (!(tc > T(-1e-8))) err = PPB_ST_ASSERT_TCOL_NEG;
                if (partner >= 0 && partner < nWl) vel = M::reflect(nrm, vel);
                else vel = fv;
              } else {
                move_ball(pos, vel, t, g);
                tc = tc - t;
                if (!(tc > T(-1e-8))) err = PPB_ST_ASSERT_TCOL_NEG;  // primitives.py:610
              }
            }
            {
              const int e = first_error<LB>(err);
              if (e) {
                status = e;
                break;
              }
            }
            tStep += t;
            ++cnt_iters;
          }
          if (status == 0 && !P.scan_only) ++s_done;
        }
      }
      // ---- what the caller wants to see of this step (a halted world keeps recording where it stopped) ----
      if (rec_traj) {
        T2 p;
        p.x = pos.x; p.y = pos.y;
        *traj_p = p;
        traj_p += total;
      }
      if (s == next_draw) {
        T2 p;
        p.x = pos.x; p.y = pos.y;
        if (has_ball && writer) *snap_p = p;
        snap_p += total;
        next_draw += P.draw_every;
      }
    }

    // ---- write back (the owning lane of every ball) ----
    if (!skip && (int)PPB_STATUS_OF(h.flags) == 0) {
      if (has_ball && writer) {
        T2 p, v;
        p.x = pos.x; p.y = pos.y; v.x = vel.x; v.y = vel.y;
        gpos[base + bi] = p;
        if (vel_dirty) gvel[base + bi] = v;
        gtcol[base + bi] = tc;
        if (rescanned || (fric && warm)) {
          T4 a;
          a.x = nrm.x; a.y = nrm.y; a.z = fv.x; a.w = fv.y;
          st4(gaux + base + bi, a);
          S.partner[base + bi] = partner;
        }
      }
      if (wt == 0 && lane == 0) {
        Hdr<T> o;
        memset(&o, 0, sizeof o);
        o.tmin = mnt;
        o.step = h.step + (uint32_t)s_done;   // a halted world keeps the index of the step it raised in
        o.flags = status ? ((uint32_t)status << 8) : (pending ? PPB_FLAG_PENDING : 0u);
        hdr[w] = o;
        if (nev_local) S.nev[w] = nev_base + nev_local;
        if (S.event_capacity <= 0 && nev_local) atomicAdd(S.counters + 0, (unsigned long long)nev_local);
      }
    }
    }   // worlds of this ticket
    if (TW == 1) {
      idx = __shfl_sync(FULL, idx_next, 0);
    } else {
      if (wt == 0 && lane == 0) ts.ticket = idx_next;
      team_sync<TW>(team);
      idx = ts.ticket;
      team_sync<TW>(team);
    }
  }
  // per-team statistics (every lane counted the same)
  if (wt == 0 && lane == 0 && (cnt_iters | cnt_rescans | cnt_slow)) {
    atomicAdd(S.counters + 1, cnt_slow);
    atomicAdd(S.counters + 2, cnt_iters);
    atomicAdd(S.counters + 3, cnt_rescans);
  }
}

// Dynamics._include_object for a range of worlds (primitives.py:556-563): tCol_ = 0, objCol_ = None,
// every ball queued; also used by add_all_dynamic_collision_queue (keep_cache = 1)
template <typename T>
__global__ void reset_cache_kernel(StatePtrs S, int n_balls, int world0, int count, int keep_cache, int32_t step_value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int w = world0 + i;
  Hdr<T> *hdr = reinterpret_cast<Hdr<T> *>(S.hdr);
  Hdr<T> h = hdr[w];
  if (keep_cache) {
    if (PPB_STATUS_OF(h.flags) == 0u) h.flags |= PPB_FLAG_PENDING;
    hdr[w] = h;
    return;
  }
  h.tmin = T(0);
  h.step = (uint32_t)step_value;
  h.flags = PPB_FLAG_PENDING;
  hdr[w] = h;
  S.nev[w] = 0u;
  typedef typename Vec4<T>::type T4;
  for (int b = 0; b < n_balls; ++b) {
    reinterpret_cast<T *>(S.tcol)[(long long)w * n_balls + b] = T(0);
    S.partner[(long long)w * n_balls + b] = -1;
    T4 z;
    z.x = z.y = z.z = z.w = T(0);
    st4(reinterpret_cast<T4 *>(S.aux) + (long long)w * n_balls + b, z);
  }
}

// Dynamics.apply_force for every ball of worlds [world0, world0 + count) (primitives.py:580-594):
//   a = (1.0 / mass) * force;  deltaV = ((0.5 * forceT) * forceT) * a;  vel = vel + deltaV
// in the reference's operation order, evaluated in fp64 whatever the batch dtype (the force array is fp64) with
// explicitly rounded multiplies / adds, so no FMA contraction can change a bit.  Collision caches are left alone,
// as in the reference (quirk Q7: only safe before the first step or together with ppb_requeue).
template <typename T>
__global__ void apply_force_kernel(StatePtrs S, int n_balls, int world0, int count, const double *force, double force_t) {
  typedef typename Vec2<T>::type T2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)count * n_balls) return;
  const long long slot = (long long)world0 * n_balls + i;
  const T2 q = reinterpret_cast<const T2 *>(S.rm)[slot];
  if (q.x < T(0)) return;   // empty slot
  const double inv_m = __ddiv_rn(1.0, (double)q.y);
  const double k = __dmul_rn(__dmul_rn(0.5, force_t), force_t);
  T2 v = reinterpret_cast<T2 *>(S.vel)[slot];
  v.x = (T)__dadd_rn((double)v.x, __dmul_rn(__dmul_rn(inv_m, force[2 * i]), k));
  v.y = (T)__dadd_rn((double)v.y, __dmul_rn(__dmul_rn(inv_m, force[2 * i + 1]), k));
  reinterpret_cast<T2 *>(S.vel)[slot] = v;
}

// Range of the radii of all occupied ball slots: out[0] = min floor(r), out[1] = max ceil(r), out[2] = 1 if some radius
// is not an integer.  The renderer draws a ball from the pre-rendered sprite of its integer radius (get_ball_im takes
// the radius as an array shape, primitives.py:39-40), so the API refuses to draw anything else (ppb_render_async).
// The same pass copies every slot's radius (fp32; integral radii are exact) into the world's render record.
template <typename T>
__global__ void radius_range_kernel(StatePtrs S, int n_walls, int n_balls, int32_t *out) {
  typedef typename Vec2<T>::type T2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int lo = 0x7fffffff, hi = -1, frac = 0;
  if (i < (long long)S.n_worlds * n_balls) {
    const T r = reinterpret_cast<const T2 *>(S.rm)[i].x;
    if (S.wall_rc) {
      const long long w = i / n_balls;
      const int b = (int)(i - w * n_balls);
      reinterpret_cast<float *>(reinterpret_cast<char *>(S.wall_rc) + (size_t)w * render_rec_stride(n_walls, n_balls) +
                                (size_t)n_walls * sizeof(WallRC))[b] = (float)r;
    }
    if (r >= T(0)) {
      const double rd = (double)r;
      lo = (int)floor(rd);
      hi = (int)ceil(rd);
      frac = floor(rd) != rd;
    }
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  frac = __any_sync(0xffffffffu, frac);
  if ((threadIdx.x & 31) == 0) {
    if (hi >= 0) {
      atomicMin(out + 0, lo);
      atomicMax(out + 1, hi);
    }
    if (frac) atomicOr(out + 2, 1);
  }
}

}  // namespace ppb
