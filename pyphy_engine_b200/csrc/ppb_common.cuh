// ppb_common.cuh -- types shared by the billiards kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pyphy_b200.h"

namespace ppb {

// ---- small 2-vector with the reference's operation order (geometry.py:9-175) ----
template <typename T>
struct V2 {
  T x, y;
};
template <typename T>
__device__ __forceinline__ V2<T> mk(T x, T y) {
  V2<T> v;
  v.x = x;
  v.y = y;
  return v;
}
template <typename T>
__device__ __forceinline__ V2<T> operator+(V2<T> a, V2<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T>
__device__ __forceinline__ V2<T> operator-(V2<T> a, V2<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T>
__device__ __forceinline__ V2<T> operator*(V2<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
template <typename T>
__device__ __forceinline__ T dot(V2<T> a, V2<T> b) { return a.x * b.x + a.y * b.y; }

__device__ __forceinline__ float ppb_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double ppb_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float ppb_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double ppb_abs(double x) { return fabs(x); }

template <typename T>
struct Lim;
template <>
struct Lim<float> {
  __device__ __forceinline__ static float inf() { return __int_as_float(0x7f800000); }
};
template <>
struct Lim<double> {
  __device__ __forceinline__ static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};

// ---- 4-wide storage vectors (16 B for fp32, 32 B for fp64) ----
template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  typedef float4 type;
};
template <>
struct Vec4<double> {
  typedef double4 type;
};
template <typename T>
struct Vec2;
template <>
struct Vec2<float> {
  typedef float2 type;
};
template <>
struct Vec2<double> {
  typedef double2 type;
};

__device__ __forceinline__ float4 ld4(const float4 *p) { return *p; }
__device__ __forceinline__ void st4(float4 *p, float4 v) { *p = v; }
__device__ __forceinline__ double4 ld4(const double4 *p) {
  // two 128-bit loads
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = q[0], b = q[1];
  double4 r;
  r.x = a.x; r.y = a.y; r.z = b.x; r.w = b.y;
  return r;
}
__device__ __forceinline__ void st4(double4 *p, double4 v) {
  double2 *q = reinterpret_cast<double2 *>(p);
  q[0] = make_double2(v.x, v.y);
  q[1] = make_double2(v.z, v.w);
}

// ---- per-world header: 16 bytes, read by every kernel ----
// flags: bit 0 = colQueue_ non-empty (every ball is rescanned at the next loop head,
// primitives.py:634-638); bits 8..15 = PPB_ST_* status (0 = running)
template <typename T>
struct Hdr;
template <>
struct __align__(16) Hdr<float> {
  float tmin;      // min over balls of tCol_ (primitives.py:641-646)
  uint32_t step;   // completed step() calls
  uint32_t flags;
  uint32_t pad;
};
template <>
struct __align__(16) Hdr<double> {
  double tmin;
  uint32_t step;
  uint32_t flags;
};
#define PPB_FLAG_PENDING 1u

#define PPB_STATUS_OF(flags) (((flags) >> 8) & 0xffu)

// layout shared by all worlds of a batch (kernel parameter, lives in constant bank)
struct Layout {
  int32_t n_balls, n_walls;
  int32_t n_obj;                                     // n_walls + n_balls
  int16_t order[PPB_MAX_WALLS + PPB_MAX_BALLS];      // insertion sequence -> object index
  int16_t rank_wall[PPB_MAX_WALLS];                  // object -> position in the sequence
  int16_t rank_ball[PPB_MAX_BALLS];
  // scan items of one ball's time_to_collide_all in visiting order (primitives.py:724-733,
  // dynamics.py:20): a wall contributes its four edges (wall << 2 | edge), a ball 0x8000 | slot
  int32_t n_items;
  int16_t item[PPB_MAX_WALLS * 4 + PPB_MAX_BALLS];
  // the inverse: position in that visiting order of edge (wall << 2 | edge) and of ball slot b
  int16_t edge_k[PPB_MAX_WALLS * 4];
  int16_t ball_k[PPB_MAX_BALLS];
};

// device state of a batch (raw pointers; T-typed views are made in the kernels)
struct StatePtrs {
  void *pos;       // [n_worlds][n_balls] {x, y}
  void *snap;      // NULL or [slots][n_worlds][n_balls] {x, y}: copies of the positions after the steps that are drawn
                   // (the render kernels read them while the physics of the following steps already runs)
  void *vel;       // [n_worlds][n_balls] {vx, vy}
  void *aux;       // [n_worlds][n_balls] {nrmlCol.x, nrmlCol.y, futureVel.x, futureVel.y}
  void *tcol;      // [n_worlds][n_balls]
  void *rm;        // [n_worlds][n_balls] {radius, mass}
  void *wall;      // [n_worlds][n_walls][4][2]
  void *wall_rc;   // [n_worlds] render records (render_rec_stride bytes: WallRC x n_walls, then the radii): what the render kernel needs of a wall, derived once per
                   // ppb_set_walls / ppb_set_styles (walls never move)
  void *hdr;       // [n_worlds] Hdr<T>
  int32_t *partner;   // [n_worlds][n_balls]
  uint32_t *nev;      // [n_worlds] events emitted so far (sequence numbers)
  ppb_event *events;  // [event_capacity]
  unsigned long long *counters;  // [0] events produced [1] world-steps that ran the event loop [2] its iterations [3] rescans
  int32_t *list;                 // (unused)
  int32_t *list_count;           // [2..3] work tickets of the advance kernel, indexed by launch parity
  long long event_capacity;
  int32_t n_worlds;
};

struct StepParams {
  double dt, gx, gy, friction;
  int32_t step_off;   // position of this step inside the running call (row of traj)
  int32_t step_abs;   // index of the step this launch advances (the batch counts its steps on the host): passed by
                      // value, so no kernel starts with a dependent load of a device-side counter
  int32_t max_substeps;
  int32_t scan_only;  // drain the collision queue (rescan every ball) of the worlds that have one and stop; nothing moves
  int32_t n_steps;    // steps this launch advances every world by
  int32_t draw_first; // index inside the launch of the first step whose positions go to a snapshot slot ...
  int32_t draw_every; // ... then every draw_every-th step (0: no snapshots); slot j lies j * n_worlds * n_balls further on
  int32_t parity;     // which work-ticket counter this launch uses (it resets the other one for the next launch)
  int32_t ticket_worlds;  // consecutive worlds a team takes per ticket
};

// ---- render ----
struct __align__(16) WallRC {
  float x0, x1, y0, y1;       // bounding box of the quad
  int bx0, bx1, by0, by1;     // integer rectangle the reference pastes the wall sprite into
  int lo, hi;                 // rows [lo, hi) the wall can touch
  int flags;                  // bit 0: axis-aligned, bit 1: axis-aligned with integral y edges
  uint32_t col;               // premultiplied RGBA8 of the wall slot
};
// The render kernel's static input of one world is ONE contiguous record: the WallRC of its walls followed by the
// radii of its ball slots as fp32 (< 0 = empty slot), padded to 16 bytes.  One region per frame instead of two keeps
// the scattered reads that interrupt the frame write stream in HBM to a minimum (each separately read array costs a
// launch of 131,072 frames ~14 us, whatever its size: measured, profiles/r2_experiments_last_session.log).
__host__ __device__ inline size_t render_rec_stride(int n_walls, int n_balls) {
  return (size_t)n_walls * sizeof(WallRC) + ((((size_t)n_balls * sizeof(float)) + 15) & ~(size_t)15);
}

struct RenderWall {  // per wall slot, shared by the batch
  int32_t kind;
  uint32_t rgba8;    // colour as premultiplied 8-bit r | g<<8 | b<<16 | a<<24 (array channel order)
};
struct RenderLayout {
  int32_t W, H;
  int32_t r_min, r_max;
  int32_t sprite_stride;                 // bytes between sprites of consecutive radii of one slot
  int32_t sprite_slot_stride;            // bytes between ball slots in the atlas
  int32_t s_thick[PPB_MAX_BALLS];
  RenderWall wall[PPB_MAX_WALLS];
};

}  // namespace ppb
