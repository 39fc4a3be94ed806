// ppb_kernels.cuh -- the per-step kernels of the batched billiards engine (sm_100a).
//
//   K1 integrate_kernel : Dynamics.move_object (primitives.py:600-611) for every ball of the
//                         worlds whose step is event-free (no queued rescan and min tCol_ > dt):
//                         a pure HBM stream, one thread per ball pair / ball.
//   K2 collide_kernel   : the event-driven step loop (primitives.py:628-733 + dynamics.py:8-136)
//                         for the remaining worlds; one warp works on one world at a time with
//                         its balls and wall vertices staged in shared memory, the
//                         (ball x object) TOC tests spread over the lanes and reduced with
//                         warp shuffles (lexicographic (t, insertion rank) arg-min).
//   K3 render_kernel    : World.generate_image (primitives.py:991-1008) -- in ppb_render.cuh.
//
// A world is advanced exactly once per step: K1 takes it when its header says the step is
// event-free and bumps header.step; K2 takes every world whose header.step still equals the
// step index.
#pragma once
#include "ppb_math.cuh"

namespace ppb {

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
// K1: integrate
// ------------------------------------------------------------------------------------
template <typename T, int BPT /* balls per thread: 1 or 2 */>
__global__ void __launch_bounds__(256) integrate_kernel(StatePtrs S, int n_balls, StepParams P, T *traj) {
  typedef typename Vec2<T>::type T2;
  const long long total = (long long)S.n_worlds * n_balls;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * BPT;
  if (i >= total) return;
  const int w = (int)(i / n_balls);
  const Hdr<T> h = reinterpret_cast<const Hdr<T> *>(S.hdr)[w];
  const uint32_t k = (uint32_t)(*S.step_base + P.step_off);
  const T dt = (T)P.dt;
  if (!(h.step == k && h.flags == 0u && h.tmin > dt)) return;
  const V2<T> g = mk<T>((T)P.gx, (T)P.gy);
  T2 *pos = reinterpret_cast<T2 *>(S.pos);
  T2 *vel = reinterpret_cast<T2 *>(S.vel);
  T *tcol = reinterpret_cast<T *>(S.tcol);
  const bool has_g = (P.gx != 0.0) || (P.gy != 0.0);
  if (BPT == 2) {
    typedef typename Vec4<T>::type T4;
    T4 p = ld4(reinterpret_cast<const T4 *>(pos + i));
    T4 v = ld4(reinterpret_cast<const T4 *>(vel + i));
    T2 tc = *reinterpret_cast<const T2 *>(tcol + i);
    V2<T> p0 = mk<T>(p.x, p.y), p1 = mk<T>(p.z, p.w), v0 = mk<T>(v.x, v.y), v1 = mk<T>(v.z, v.w);
    move_ball(p0, v0, dt, g);
    move_ball(p1, v1, dt, g);
    tc.x = tc.x - dt;
    tc.y = tc.y - dt;
    T4 po;
    po.x = p0.x; po.y = p0.y; po.z = p1.x; po.w = p1.y;
    st4(reinterpret_cast<T4 *>(pos + i), po);
    if (has_g) {
      T4 vo;
      vo.x = v0.x; vo.y = v0.y; vo.z = v1.x; vo.w = v1.y;
      st4(reinterpret_cast<T4 *>(vel + i), vo);
    }
    *reinterpret_cast<T2 *>(tcol + i) = tc;
    if (traj) st4(reinterpret_cast<T4 *>(traj + ((long long)P.step_off * total + i) * 2), po);
  } else {
    T2 p = pos[i], v = vel[i];
    T tc = tcol[i];
    V2<T> p0 = mk<T>(p.x, p.y), v0 = mk<T>(v.x, v.y);
    move_ball(p0, v0, dt, g);
    tc = tc - dt;
    p.x = p0.x; p.y = p0.y;
    pos[i] = p;
    if (has_g) {
      v.x = v0.x; v.y = v0.y;
      vel[i] = v;
    }
    tcol[i] = tc;
    if (traj) reinterpret_cast<T2 *>(traj)[(long long)P.step_off * total + i] = p;
  }
  if (i - (long long)w * n_balls == 0) {
    Hdr<T> o = h;
    o.tmin = h.tmin - dt;   // min_b (tCol_b - dt) == (min_b tCol_b) - dt: rounding is monotone
    o.step = k + 1u;
    reinterpret_cast<Hdr<T> *>(S.hdr)[w] = o;
  }
}

// ------------------------------------------------------------------------------------
// K2: collide
// ------------------------------------------------------------------------------------
template <typename T>
struct WarpSmem {
  typename Vec2<T>::type pos[PPB_MAX_BALLS];
  typename Vec2<T>::type vel[PPB_MAX_BALLS];
  typename Vec2<T>::type rm[PPB_MAX_BALLS];
  T wall[PPB_MAX_WALLS * 8];
};

#define PPB_COLLIDE_WARPS 4

template <typename T>
__global__ void __launch_bounds__(PPB_COLLIDE_WARPS * 32)
collide_kernel(StatePtrs S, Layout L, StepParams P, T *traj) {
  typedef typename MathOf<T>::type M;
  typedef typename Vec2<T>::type T2;
  typedef typename Vec4<T>::type T4;
  __shared__ WarpSmem<T> smem[PPB_COLLIDE_WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  WarpSmem<T> &sm = smem[warp];
  const unsigned FULL = 0xffffffffu;
  const int nB = L.n_balls, nWl = L.n_walls;
  const uint32_t k = (uint32_t)(*S.step_base + P.step_off);
  const T dt = (T)P.dt;
  const V2<T> g = mk<T>((T)P.gx, (T)P.gy);
  const T INF = Lim<T>::inf();

  // lanes per ball: G = 32 / (next power of two >= n_balls)
  int nbp2 = 1;
  while (nbp2 < nB) nbp2 <<= 1;
  const int G = 32 / nbp2;
  const int bi = lane / G;     // ball handled by this lane group
  const int sub = lane - bi * G;
  const bool has_ball = bi < nB;
  const bool owner = has_ball && sub == 0;
  const int J = nWl * 4 + nB;  // objects a ball is tested against (edges, then balls)

  Hdr<T> *hdr = reinterpret_cast<Hdr<T> *>(S.hdr);
  T2 *gpos = reinterpret_cast<T2 *>(S.pos);
  T2 *gvel = reinterpret_cast<T2 *>(S.vel);
  T2 *grm = reinterpret_cast<T2 *>(S.rm);
  T4 *gaux = reinterpret_cast<T4 *>(S.aux);
  T *gtcol = reinterpret_cast<T *>(S.tcol);
  const T *gwall = reinterpret_cast<const T *>(S.wall);

  const int n_groups = (S.n_worlds + 31) >> 5;
  unsigned long long cnt_worlds = 0, cnt_iters = 0, cnt_rescans = 0;

  for (int grp = blockIdx.x * PPB_COLLIDE_WARPS + warp; grp < n_groups; grp += gridDim.x * PPB_COLLIDE_WARPS) {
    const int wl = (grp << 5) + lane;
    Hdr<T> h;
    h.tmin = INF;
    h.step = 0xffffffffu;
    h.flags = 0;
    if (wl < S.n_worlds) h = hdr[wl];
    unsigned todo = __ballot_sync(FULL, wl < S.n_worlds && (h.step == k || PPB_STATUS_OF(h.flags) != 0u));
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int w = (grp << 5) + src;
      const uint32_t flags = __shfl_sync(FULL, h.flags, src);
      const long long base = (long long)w * nB;

      if (PPB_STATUS_OF(flags) != 0u) {
        // halted world: nothing moves any more; keep recording where it stopped
        if (traj && lane < nB)
          reinterpret_cast<T2 *>(traj)[((long long)P.step_off * S.n_worlds + w) * nB + lane] = gpos[base + lane];
        continue;
      }

      // ---- stage the world ----
      V2<T> pos = mk<T>(0, 0), vel = pos, nrm = pos, fv = pos;
      T tc = INF, rad = T(-1), mass = T(1);
      int partner = -1;
      if (has_ball) {
        T2 p = gpos[base + bi], v = gvel[base + bi], q = grm[base + bi];
        T4 a = ld4(gaux + base + bi);
        pos = mk<T>(p.x, p.y);
        vel = mk<T>(v.x, v.y);
        rad = q.x;
        mass = q.y;
        nrm = mk<T>(a.x, a.y);
        fv = mk<T>(a.z, a.w);
        tc = gtcol[base + bi];
        partner = S.partner[base + bi];
        if (sub == 0) {
          sm.pos[bi] = p;
          sm.vel[bi] = v;
          sm.rm[bi] = q;
        }
      }
      for (int j = lane; j < nWl * 8; j += 32) sm.wall[j] = gwall[(long long)w * nWl * 8 + j];
      __syncwarp();
      const bool active = has_ball && rad >= T(0);

      bool pending = (flags & PPB_FLAG_PENDING) != 0u;
      bool rescanned = false;
      int err = 0;
      T tStep = T(0), mnt = INF;
      uint32_t nev_local = 0;
      int iters = 0;

      for (;;) {
        if (pending) {  // colQueue_ drain: time_to_collide_all for every ball (primitives.py:634-638, 720-733)
          pending = false;
          rescanned = true;
          ++cnt_rescans;
          T bt = INF;
          int bkey = 0x7fffffff, bpart = -1;
          V2<T> bn = mk<T>(0, 0);
          int fkey = -1;
          V2<T> bfv = fv;
          int ekey = 0x7fffffff;
          if (active) {
            for (int j = sub; j < J; j += G) {
              Toc<T> r;
              int key, part;
              if (j < nWl * 4) {
                const int q = j >> 2, e = j & 3, e1 = (e + 1) & 3;
                const T *wv = sm.wall + q * 8;
                r = M::edge(pos, vel, rad, mk<T>(wv[2 * e], wv[2 * e + 1]), mk<T>(wv[2 * e1], wv[2 * e1 + 1]));
                key = ((int)L.rank_wall[q] << 2) | e;
                part = q;
              } else {
                const int b2 = j - nWl * 4;
                const T2 q2 = sm.rm[b2];
                if (b2 == bi || q2.x < T(0)) continue;
                const T2 p2 = sm.pos[b2], v2 = sm.vel[b2];
                r = M::ball(pos, vel, rad, mass, mk<T>(p2.x, p2.y), mk<T>(v2.x, v2.y), q2.x, q2.y);
                key = (int)L.rank_ball[b2] << 2;
                part = nWl + b2;
              }
              if (r.err) {
                const int ek = (key << 8) | r.err;
                ekey = ek < ekey ? ek : ekey;
                continue;
              }
              if (r.has_fv && key > fkey) {  // futureVel_: the last partner in world order wins (dynamics.py:129)
                fkey = key;
                bfv = r.fv;
              }
              if (r.t < bt || (r.t == bt && key < bkey)) {  // strict '<' in insertion order (primitives.py:729)
                bt = r.t;
                bkey = key;
                bpart = part;
                bn = r.n;
              }
            }
          }
          // reduce over the G lanes of this ball
          for (int o = G >> 1; o > 0; o >>= 1) {
            const T ot = __shfl_xor_sync(FULL, bt, o);
            const int ok = __shfl_xor_sync(FULL, bkey, o);
            const int op = __shfl_xor_sync(FULL, bpart, o);
            const T onx = __shfl_xor_sync(FULL, bn.x, o), ony = __shfl_xor_sync(FULL, bn.y, o);
            const int ofk = __shfl_xor_sync(FULL, fkey, o);
            const T ofx = __shfl_xor_sync(FULL, bfv.x, o), ofy = __shfl_xor_sync(FULL, bfv.y, o);
            const int oek = __shfl_xor_sync(FULL, ekey, o);
            if (ot < bt || (ot == bt && ok < bkey)) {
              bt = ot; bkey = ok; bpart = op; bn = mk<T>(onx, ony);
            }
            if (ofk > fkey) {
              fkey = ofk; bfv = mk<T>(ofx, ofy);
            }
            ekey = oek < ekey ? oek : ekey;
          }
          // first error in (ball, object) order is the one the reference raises
          unsigned long long e64 = (ekey == 0x7fffffff || !has_ball)
                                       ? 0xffffffffffffffffull
                                       : (((unsigned long long)bi << 40) | (unsigned long long)(unsigned)ekey);
          e64 = warp_min(e64);
          if (e64 != 0xffffffffffffffffull) {
            err = (int)(e64 & 0xffull);
            break;
          }
          if (active) {
            tc = bt;
            partner = bpart;
            if (bpart >= 0) nrm = bn;
            fv = bfv;
          }
        }
        // mntCol / t  (primitives.py:641-650)
        mnt = warp_min(active ? tc : INF);
        const T rem = dt - tStep;
        const T t = (rem < mnt) ? rem : mnt;
        if (t <= T(0)) break;
        if (++iters > P.max_substeps) {
          err = PPB_ST_ITER_LIMIT;
          break;
        }
        const bool hit = active && (tc <= t);  // step_object, primitives.py:617
        const unsigned evm = __ballot_sync(FULL, hit && sub == 0);
        if (evm) {
          // one record per resolve_collision call, in ball order
          unsigned long long slot0 = 0;
          const int n_ev = __popc(evm);
          if (lane == 0 && S.event_capacity > 0) slot0 = atomicAdd(S.counters + 0, (unsigned long long)n_ev);
          slot0 = __shfl_sync(FULL, slot0, 0);
          if (hit && sub == 0) {
            const int rnk = __popc(evm & ((1u << lane) - 1u));
            if (S.event_capacity > 0 && (long long)(slot0 + rnk) < S.event_capacity) {
              ppb_event ev;
              ev.world = w;
              ev.seq = (int32_t)(S.nev[w] + nev_local + rnk);
              ev.step = (int32_t)k;
              ev.ball = (int16_t)bi;
              ev.partner = (int16_t)partner;
              S.events[slot0 + rnk] = ev;
            }
          }
          nev_local += n_ev;
          pending = true;  // add_all_dynamic_collision_queue, primitives.py:697-700
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
          if (sub == 0) {
            T2 p, v;
            p.x = pos.x; p.y = pos.y; v.x = vel.x; v.y = vel.y;
            sm.pos[bi] = p;
            sm.vel[bi] = v;
          }
        }
        __syncwarp();
        {  // the first ball (in order) whose move raised decides the status
          unsigned em = __ballot_sync(FULL, err != 0 && sub == 0);
          if (em) {
            err = __shfl_sync(FULL, err, __ffs(em) - 1);
            break;
          }
        }
        tStep += t;
        ++cnt_iters;
      }

      // ---- write back ----
      if (owner) {
        T2 p, v;
        p.x = pos.x; p.y = pos.y; v.x = vel.x; v.y = vel.y;
        gpos[base + bi] = p;
        gvel[base + bi] = v;
        gtcol[base + bi] = tc;
        if (rescanned) {
          T4 a;
          a.x = nrm.x; a.y = nrm.y; a.z = fv.x; a.w = fv.y;
          st4(gaux + base + bi, a);
          S.partner[base + bi] = partner;
        }
        if (traj) reinterpret_cast<T2 *>(traj)[((long long)P.step_off * S.n_worlds + w) * nB + bi] = p;
      }
      if (lane == 0) {
        Hdr<T> o;
        o.tmin = mnt;
        o.step = err ? k : k + 1u;  // a halted world keeps the index of the step it raised in
        o.flags = err ? ((uint32_t)err << 8) : (pending ? PPB_FLAG_PENDING : 0u);
        hdr[w] = o;
        if (nev_local) S.nev[w] += nev_local;
        if (S.event_capacity <= 0 && nev_local) atomicAdd(S.counters + 0, (unsigned long long)nev_local);
      }
      ++cnt_worlds;
      __syncwarp();
    }
  }
  if (lane == 0 && cnt_worlds) {
    atomicAdd(S.counters + 1, cnt_worlds);
    atomicAdd(S.counters + 2, cnt_iters);
    atomicAdd(S.counters + 3, cnt_rescans);
  }
}

// advance the device-side step counter after a group of step launches
static __global__ void advance_step_base_kernel(int32_t *step_base, int32_t n) { *step_base += n; }

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

}  // namespace ppb
