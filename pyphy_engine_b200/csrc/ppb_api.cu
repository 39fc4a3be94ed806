// ppb_api.cu -- C ABI of the batched billiards engine (include/pyphy_b200.h).
// Host side only: owns the device state, converts host fp64 arrays to the batch dtype,
// sequences the K1/K2/K3 launches of a step and the copy pipeline of ppb_simulate_host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "ppb_launch.h"
#include "ppb_pack.h"

using namespace ppb;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(PPB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                                       __FILE__, __LINE__);                                       \
  } while (0)

inline uint32_t h_color_to_u8(double v) {
  if (v < 0) v = 0;
  if (v > 1) v = 1;
  return ((uint32_t)(v * 65535.0 + 0.5)) >> 8;
}
inline uint32_t h_premul_rgba8(const float c[4]) {
  const double a = c[3];
  return h_color_to_u8(c[0] * a) | (h_color_to_u8(c[1] * a) << 8) | (h_color_to_u8(c[2] * a) << 16) |
         (h_color_to_u8(a) << 24);
}

}  // namespace

// Steps per advance launch: long enough to amortise the state round trip and to let the warps that own a world with
// a long sub-step chain catch up, short enough that a launch stays in the millisecond range.  With frames, a launch
// covers at most PPB_DRAWS_PER_CHUNK drawn steps (their position snapshots are what the render kernels read).
#define PPB_STEPS_PER_LAUNCH 64
#define PPB_TICKET_WORLDS_DEFAULT 2
#define PPB_DRAWS_PER_CHUNK 64

struct ppb_batch {
  ppb_config cfg;
  Layout L;
  StatePtrs S;
  StepParams P;
  RenderLayout RL;
  int sm_count = 0;
  int advance_grid = 1;
  int launch_parity = 0;   // work-ticket counter the next advance launch uses
  size_t esz = 4;
  bool styles_set = false;
  uint32_t *atlas = nullptr;
  size_t atlas_bytes = 0;
  uint32_t *d_fill8 = nullptr, *d_stroke8 = nullptr;
  int64_t launches = 0;
  int32_t host_step = 0;  // index of the next step; every launch receives it by value
  // profiling (CUDA events around every kernel of the last ppb_step_async)
  int profiling = 0;   // bit k: CUDA events around kernel kind k (0 integrate, 1 collide, 2 render)
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_kind;  // kernel kind per (start, stop) pair
  size_t ev_used = 0;
  // software pipeline of ppb_step_async: render of step k on its own stream, overlapped with the physics of step k+1
  cudaStream_t render_stream = nullptr;
  void *snap[2] = {nullptr, nullptr};   // PPB_DRAWS_PER_CHUNK position snapshots each
  cudaEvent_t ev_phys[2] = {nullptr, nullptr}, ev_rend[2] = {nullptr, nullptr}, ev_gate[2] = {nullptr, nullptr};
  int draws_per_chunk = PPB_DRAWS_PER_CHUNK;
  int overlap_renders = 2;   // how many of a chunk's last render launches the next chunk's physics may run beside (-1: all)
  // page-locked staging for host <-> device state transfers (grown on demand)
  void *stage = nullptr;
  size_t stage_bytes = 0;
  // scratch of ppb_simulate_host
  void *scr_traj[2] = {nullptr, nullptr};
  float *scr_traj32[2] = {nullptr, nullptr};
  uint8_t *scr_frames[2] = {nullptr, nullptr};
  size_t scr_traj_bytes = 0, scr_frames_bytes = 0;
  int scr_chunk = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  unsigned sim_it = 0;   // staging buffers used so far (they alternate across ppb_simulate_host* calls)
  cudaEvent_t ev_call[2] = {nullptr, nullptr};   // end of the copies of the last two ppb_simulate_host* calls
  unsigned sim_calls = 0;
  // frames of ppb_simulate_host* travel run-length packed and are rebuilt by host threads (ppb_pack.cu)
  PackPipe *pack = nullptr;
  bool pack_frames = true;
  long long pack_mark[2] = {0, 0};   // jobs submitted before each of the last two ppb_simulate_host* calls
  // radii of the occupied ball slots, measured on the device the first time a frame is requested after the balls
  // changed: {min floor(r), max ceil(r), any non-integer}
  int32_t *d_rchk = nullptr;
  bool radii_known = false;
  int32_t rad_lo = 0, rad_hi = -1, rad_frac = 0;
};

namespace {

bool is64(const ppb_batch *b) { return b->cfg.dtype == PPB_F64; }

int check_range(const ppb_batch *b, int32_t world0, int32_t count) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (world0 < 0 || count < 0 || (int64_t)world0 + count > b->cfg.n_worlds)
    return fail(PPB_E_INVALID, "world range [%d, %d) outside [0, %d)", world0, world0 + count, b->cfg.n_worlds);
  return PPB_OK;
}

// page-locked staging buffer of at least `bytes`
int stage_reserve(ppb_batch *b, size_t bytes) {
  if (bytes <= b->stage_bytes) return PPB_OK;
  if (b->stage) cudaFreeHost(b->stage);
  b->stage = nullptr;
  b->stage_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  if (cudaHostAlloc(&b->stage, want, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return fail(PPB_E_NOMEM, "cudaHostAlloc(%zu bytes) failed", want);
  }
  b->stage_bytes = want;
  return PPB_OK;
}

// element-wise conversion spread over a few host threads (large batches only)
template <typename D, typename S>
void convert_array(D *dst, const S *src, size_t n) {
  const size_t kMin = 1u << 18;
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 8) nt = 8;
  if (n < 2 * kMin || nt < 2) {
    for (size_t i = 0; i < n; ++i) dst[i] = (D)src[i];
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const size_t lo = t * per, hi = (lo + per < n) ? lo + per : n;
    if (lo >= hi) break;
    th.emplace_back([=]() {
      for (size_t i = lo; i < hi; ++i) dst[i] = (D)src[i];
    });
  }
  for (auto &t : th) t.join();
}

// host fp64 -> device array of the batch dtype, through the page-locked staging buffer
int upload(ppb_batch *b, void *dst_base, size_t elem_off, const double *src, size_t n, cudaStream_t st) {
  if (n == 0) return PPB_OK;
  int rc = stage_reserve(b, n * b->esz);
  if (rc) return rc;
  if (is64(b)) memcpy(b->stage, src, n * 8);
  else convert_array((float *)b->stage, src, n);
  CU(cudaMemcpyAsync((char *)dst_base + elem_off * b->esz, b->stage, n * b->esz, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));   // the staging buffer is reused by the next call
  return PPB_OK;
}
int download(ppb_batch *b, const void *src_base, size_t elem_off, double *dst, size_t n, cudaStream_t st) {
  if (n == 0) return PPB_OK;
  int rc = stage_reserve(b, n * b->esz);
  if (rc) return rc;
  CU(cudaMemcpyAsync(b->stage, (const char *)src_base + elem_off * b->esz, n * b->esz, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (is64(b)) memcpy(dst, b->stage, n * 8);
  else convert_array(dst, (const float *)b->stage, n);
  return PPB_OK;
}

cudaError_t do_reset(ppb_batch *b, int world0, int count, int keep, cudaStream_t st) {
  ++b->launches;
  return is64(b) ? Launch<double>::reset_cache(b->S, b->cfg.n_balls, world0, count, keep, b->host_step, st)
                 : Launch<float>::reset_cache(b->S, b->cfg.n_balls, world0, count, keep, b->host_step, st);
}

struct Prof {
  ppb_batch *b;
  cudaStream_t st;
  int kind;
  bool on;
  size_t idx = 0;
  Prof(ppb_batch *b_, cudaStream_t st_, int kind_) : b(b_), st(st_), kind(kind_), on(((b_->profiling >> kind_) & 1) != 0) {
    if (!on) return;
    if (b->ev_used + 2 > b->ev_pool.size()) {
      for (int i = 0; i < 2; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) {
          on = false;
          return;
        }
        b->ev_pool.push_back(e);
      }
    }
    idx = b->ev_used;
    b->ev_used += 2;
    b->ev_kind.resize(b->ev_used / 2);
    b->ev_kind[idx / 2] = kind;
    cudaEventRecord(b->ev_pool[idx], st);
  }
  ~Prof() {
    if (on) cudaEventRecord(b->ev_pool[idx + 1], st);
  }
};

// n_steps x Dynamics.step for every world, one launch.  Positions after step draw_first, draw_first + draw_every, ...
// (indices inside the launch) go to consecutive snapshot slots of `snap` when it is given.
int step_physics(ppb_batch *b, int step_off, int n_steps, void *traj, void *snap, int draw_first, int draw_every,
                 cudaStream_t st) {
  if (n_steps <= 0) return PPB_OK;
  StepParams P = b->P;
  P.step_off = step_off;
  P.step_abs = b->host_step + step_off;
  P.n_steps = n_steps;
  P.draw_first = draw_first;
  P.draw_every = snap ? draw_every : 0;
  P.parity = b->launch_parity;
  b->launch_parity ^= 1;
  StatePtrs S = b->S;
  S.snap = snap;
  const int grid = b->advance_grid;
  Prof p(b, st, 0);
  CU(is64(b) ? Launch<double>::advance(S, b->L, P, traj, grid, st)
             : Launch<float>::advance(S, b->L, P, traj, grid, st));
  ++b->launches;
  return PPB_OK;
}

// World.generate_image() from the positions in `pos_src` (the live array or a snapshot)
int step_render(ppb_batch *b, void *pos_src, uint8_t *frame, cudaStream_t st) {
  StatePtrs S = b->S;
  S.pos = pos_src;
  Prof p(b, st, 2);
  CU(is64(b) ? Launch<double>::render(S, b->L, b->RL, b->atlas, frame, st)
             : Launch<float>::render(S, b->L, b->RL, b->atlas, frame, st));
  ++b->launches;
  return PPB_OK;
}

// A ball is drawn from the pre-rendered sprite of its integer radius; anything else is refused instead of being
// dropped or truncated silently.  The measurement (one small kernel + a 12-byte read-back) runs once after the
// balls changed, then the answer is cached.
int check_radii(ppb_batch *b, cudaStream_t st) {
  if (!b->radii_known) {
    if (!b->d_rchk) CU(cudaMalloc((void **)&b->d_rchk, 4 * sizeof(int32_t)));
    const int32_t init[4] = {0x7fffffff, -1, 0, 0};
    int32_t got[4];
    CU(cudaMemcpyAsync(b->d_rchk, init, sizeof init, cudaMemcpyHostToDevice, st));
    CU(is64(b) ? Launch<double>::radius_range(b->S, b->cfg.n_walls, b->cfg.n_balls, b->d_rchk, st)
               : Launch<float>::radius_range(b->S, b->cfg.n_walls, b->cfg.n_balls, b->d_rchk, st));
    ++b->launches;
    CU(cudaMemcpyAsync(got, b->d_rchk, sizeof got, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    b->rad_lo = got[0]; b->rad_hi = got[1]; b->rad_frac = got[2];
    b->radii_known = true;
  }
  if (b->rad_hi < 0) return PPB_OK;   // no ball anywhere
  if (b->rad_frac)
    return fail(PPB_E_STATE, "a ball radius is not an integer: the reference uses it as an array shape (get_ball_im, "
                             "primitives.py:39-40), so only integral radii can be drawn");
  if (b->rad_lo < b->RL.r_min || b->rad_hi > b->RL.r_max)
    return fail(PPB_E_STATE, "ball radii span [%d, %d] but ppb_set_styles built sprites for [%d, %d] only", b->rad_lo,
                b->rad_hi, b->RL.r_min, b->RL.r_max);
  return PPB_OK;
}

int pipeline_reserve(ppb_batch *b) {
  if (b->render_stream) return PPB_OK;
  // two sets of position snapshots, at most 1 GiB each
  const size_t one = (size_t)b->cfg.n_worlds * b->cfg.n_balls * 2 * b->esz;
  if (one && (size_t)b->draws_per_chunk * one > ((size_t)1 << 30)) b->draws_per_chunk = (int)std::max<size_t>(1, ((size_t)1 << 30) / one);
  const size_t bytes = one * b->draws_per_chunk;
  for (int i = 0; i < 2; ++i) {
    CU(cudaMalloc(&b->snap[i], bytes ? bytes : 16));
    CU(cudaEventCreateWithFlags(&b->ev_phys[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_rend[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&b->ev_gate[i], cudaEventDisableTiming));
  }
  // highest priority: the render of step k and the integrate kernel of step k+1 become ready at the same moment;
  // the renderer's one large CTA per SM should be placed before the small physics CTAs fill the SM (it does not
  // fit next to more than two collide CTAs).  Measured neutral on the 131,072 x 8 workload (0.4337 vs 0.4345 ms).
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CU(cudaStreamCreateWithPriority(&b->render_stream, cudaStreamNonBlocking, prio_hi));
  return PPB_OK;
}

void free_scratch(ppb_batch *b) {
  for (int i = 0; i < 2; ++i) {
    if (b->scr_traj[i]) cudaFree(b->scr_traj[i]);
    if (b->scr_traj32[i]) cudaFree(b->scr_traj32[i]);
    if (b->scr_frames[i]) cudaFree(b->scr_frames[i]);
    b->scr_traj[i] = nullptr;
    b->scr_traj32[i] = nullptr;
    b->scr_frames[i] = nullptr;
    if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]);
    if (b->ev_copied[i]) cudaEventDestroy(b->ev_copied[i]);
    if (b->ev_call[i]) cudaEventDestroy(b->ev_call[i]);
    b->ev_done[i] = b->ev_copied[i] = b->ev_call[i] = nullptr;
  }
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  b->copy_stream = nullptr;
  b->scr_traj_bytes = b->scr_frames_bytes = 0;
  b->scr_chunk = 0;
  b->sim_it = 0;
  b->sim_calls = 0;
}

}  // namespace

extern "C" {

const char *ppb_last_error(void) { return g_err.c_str(); }
int ppb_version(void) { return PPB_VERSION; }

int ppb_create(const ppb_config *cfg, ppb_batch **out) {
  if (!cfg || !out) return fail(PPB_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_worlds < 1) return fail(PPB_E_INVALID, "n_worlds must be >= 1");
  if (cfg->n_balls < 1 || cfg->n_balls > PPB_MAX_BALLS)
    return fail(PPB_E_INVALID, "n_balls must be in [1, %d]", PPB_MAX_BALLS);
  if (cfg->n_walls < 0 || cfg->n_walls > PPB_MAX_WALLS)
    return fail(PPB_E_INVALID, "n_walls must be in [0, %d]", PPB_MAX_WALLS);
  if (cfg->dtype != PPB_F32 && cfg->dtype != PPB_F64) return fail(PPB_E_INVALID, "dtype must be PPB_F32 or PPB_F64");
  if (!(cfg->dt > 0)) return fail(PPB_E_INVALID, "dt must be positive");
  if (cfg->canvas_w < 0 || cfg->canvas_h < 0) return fail(PPB_E_INVALID, "canvas size must be >= 0");
  if (cfg->event_capacity < 0) return fail(PPB_E_INVALID, "event_capacity must be >= 0");
  if (!(cfg->friction >= 0)) return fail(PPB_E_INVALID, "friction must be >= 0");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return fail(PPB_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PPB_E_INVALID, "device %d out of range", cfg->device);
  CU(cudaSetDevice(cfg->device));

  ppb_batch *b = new (std::nothrow) ppb_batch();
  if (!b) return fail(PPB_E_NOMEM, "out of host memory");
  b->cfg = *cfg;
  b->cfg.order = nullptr;
  const int nB = cfg->n_balls, nW = cfg->n_walls, nO = nB + nW;
  memset(&b->L, 0, sizeof b->L);
  b->L.n_balls = nB;
  b->L.n_walls = nW;
  b->L.n_obj = nO;
  std::vector<int> seen(nO, 0);
  for (int k = 0; k < nO; ++k) {
    const int q = cfg->order ? cfg->order[k] : k;
    if (q < 0 || q >= nO || seen[q]) {
      delete b;
      return fail(PPB_E_INVALID, "order must be a permutation of 0..%d", nO - 1);
    }
    seen[q] = 1;
    b->L.order[k] = (int16_t)q;
    if (q < nW) b->L.rank_wall[q] = (int16_t)k;
    else b->L.rank_ball[q - nW] = (int16_t)k;
  }
  b->L.n_items = 0;
  for (int k = 0; k < nO; ++k) {
    const int q = b->L.order[k];
    if (q < nW)
      for (int e = 0; e < 4; ++e) {
        b->L.edge_k[(q << 2) | e] = (int16_t)b->L.n_items;
        b->L.item[b->L.n_items++] = (int16_t)((q << 2) | e);
      }
    else {
      b->L.ball_k[q - nW] = (int16_t)b->L.n_items;
      b->L.item[b->L.n_items++] = (int16_t)(0x8000 | (q - nW));
    }
  }
  // Dynamics.step walks the dynamic objects in ball insertion order (primitives.py:642, 653);
  // slot numbers are that order, so ball ranks must be increasing
  for (int i = 1; i < nB; ++i)
    if (b->L.rank_ball[i] < b->L.rank_ball[i - 1]) {
      delete b;
      return fail(PPB_E_INVALID, "ball slots must appear in increasing order in `order`");
    }
  b->P.dt = cfg->dt;
  b->P.gx = cfg->gx;
  b->P.gy = cfg->gy;
  b->P.friction = cfg->friction;
  b->P.step_off = 0;
  b->P.step_abs = 0;
  b->P.scan_only = 0;
  b->P.ticket_worlds = PPB_TICKET_WORLDS_DEFAULT;
  if (const char *e = getenv("PPB_TICKET")) { const int v = atoi(e); if (v >= 1 && v <= 1024) b->P.ticket_worlds = v; }   // developer switch
  b->P.max_substeps = cfg->max_substeps > 0 ? cfg->max_substeps : (1 << 20);
  b->esz = is64(b) ? 8 : 4;
  CU(cudaDeviceGetAttribute(&b->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));

  const size_t nw = (size_t)cfg->n_worlds, nb = nw * nB;
  memset(&b->S, 0, sizeof b->S);
  b->S.n_worlds = cfg->n_worlds;
  b->S.event_capacity = cfg->event_capacity;
  struct {
    void **p;
    size_t bytes;
  } allocs[] = {
      {&b->S.pos, nb * 2 * b->esz},   {&b->S.vel, nb * 2 * b->esz},
      {&b->S.aux, nb * 4 * b->esz},   {&b->S.tcol, nb * b->esz},
      {&b->S.rm, nb * 2 * b->esz},    {&b->S.wall, nw * (size_t)(nW > 0 ? nW : 1) * 8 * b->esz},
      {&b->S.wall_rc, nw * render_rec_stride(nW, nB) + 16},
      {&b->S.hdr, nw * 16},           {(void **)&b->S.partner, nb * 4},
      {(void **)&b->S.nev, nw * 4},   {(void **)&b->S.events, (size_t)(cfg->event_capacity > 0 ? cfg->event_capacity : 1) * sizeof(ppb_event)},
      {(void **)&b->S.counters, 8 * sizeof(unsigned long long)},
      {(void **)&b->S.list, 16},
      {(void **)&b->S.list_count, 16},
  };
  for (auto &a : allocs) {
    cudaError_t e = cudaMalloc(a.p, a.bytes ? a.bytes : 16);
    if (e != cudaSuccess) {
      ppb_destroy(b);
      return fail(e == cudaErrorMemoryAllocation ? PPB_E_NOMEM : PPB_E_CUDA, "cudaMalloc(%zu bytes): %s", a.bytes,
                  cudaGetErrorString(e));
    }
    cudaMemset(*a.p, 0, a.bytes ? a.bytes : 16);
  }
  // empty worlds: every slot empty (radius -1), caches reset, all balls queued
  {
    std::vector<double> rm(nb * 2);
    for (size_t i = 0; i < nb; ++i) {
      rm[2 * i] = -1.0;
      rm[2 * i + 1] = 1.0;
    }
    int rc = upload(b, b->S.rm, 0, rm.data(), rm.size(), 0);
    if (rc != PPB_OK) {
      ppb_destroy(b);
      return rc;
    }
    cudaError_t e = do_reset(b, 0, cfg->n_worlds, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      ppb_destroy(b);
      return fail(PPB_E_CUDA, "initial reset failed: %s", cudaGetErrorString(e));
    }
  }
  b->advance_grid = is64(b) ? Launch<double>::advance_grid(cfg->n_worlds, nB, b->sm_count)
                            : Launch<float>::advance_grid(cfg->n_worlds, nB, b->sm_count);
  if (const char *e = getenv("PPB_RAW_D2H")) b->pack_frames = !(e[0] == '1');   // developer switch: plain frame copies
  if (const char *e = getenv("PPB_OVERLAP")) b->overlap_renders = atoi(e);   // developer switch (A/B timing)
  if (const char *e = getenv("PPB_CHUNK")) {
    const int v = atoi(e);
    if (v >= 1 && v <= PPB_STEPS_PER_LAUNCH) b->draws_per_chunk = v;
  }
  memset(&b->RL, 0, sizeof b->RL);
  b->RL.W = cfg->canvas_w;
  b->RL.H = cfg->canvas_h;
  *out = b;
  return PPB_OK;
}

int ppb_destroy(ppb_batch *b) {
  if (!b) return PPB_OK;
  cudaSetDevice(b->cfg.device);
  cudaDeviceSynchronize();
  if (b->pack) packpipe_destroy(b->pack);   // waits for the frame downloads in flight
  b->pack = nullptr;
  free_scratch(b);
  if (b->stage) cudaFreeHost(b->stage);
  b->stage = nullptr;
  for (int i = 0; i < 2; ++i) {
    if (b->snap[i]) cudaFree(b->snap[i]);
    if (b->ev_phys[i]) cudaEventDestroy(b->ev_phys[i]);
    if (b->ev_rend[i]) cudaEventDestroy(b->ev_rend[i]);
    if (b->ev_gate[i]) cudaEventDestroy(b->ev_gate[i]);
  }
  if (b->render_stream) cudaStreamDestroy(b->render_stream);
  void *ptrs[] = {b->S.pos, b->S.vel, b->S.aux, b->S.tcol, b->S.rm, b->S.wall, b->S.wall_rc, b->S.hdr, b->S.partner,
                  b->S.nev, b->S.events, b->S.counters, b->S.list, b->S.list_count, b->atlas, b->d_fill8,
                  b->d_stroke8, b->d_rchk};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (cudaEvent_t e : b->ev_pool) cudaEventDestroy(e);
  delete b;
  return PPB_OK;
}

int ppb_set_walls(ppb_batch *b, int32_t world0, int32_t count, const double *wall, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  if (!wall && b->cfg.n_walls > 0) return fail(PPB_E_INVALID, "wall is null");
  CU(cudaSetDevice(b->cfg.device));
  const size_t per = (size_t)b->cfg.n_walls * 8;
  rc = upload(b, b->S.wall, (size_t)world0 * per, wall, (size_t)count * per, (cudaStream_t)stream);
  if (rc) return rc;
  CU(is64(b) ? Launch<double>::wall_cache(b->S, b->L, b->RL, world0, count, (cudaStream_t)stream)
             : Launch<float>::wall_cache(b->S, b->L, b->RL, world0, count, (cudaStream_t)stream));
  ++b->launches;
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return PPB_OK;
}

int ppb_set_balls(ppb_batch *b, int32_t world0, int32_t count, const double *pos, const double *vel,
                  const double *rad, const double *mass, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  if (!pos || !vel || !rad || !mass) return fail(PPB_E_INVALID, "null array");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  // One pass over the slots, spread over host threads, straight into the page-locked staging buffer
  // (batch dtype): [pos | vel | radius,mass].  Empty slots carry zero position and velocity so that the
  // integrate kernel may stream over them.
  const size_t esz = b->esz;
  if ((rc = stage_reserve(b, n * 6 * esz))) return rc;
  char *sp = (char *)b->stage, *sv = sp + n * 2 * esz, *sr = sv + n * 2 * esz;
  const bool f64 = is64(b);
  unsigned nt = std::thread::hardware_concurrency();
  if (nt > 8) nt = 8;
  if (nt < 1 || n < (1u << 16)) nt = 1;
  std::vector<int> bad(nt, 0);
  auto work = [&](unsigned t) {
    const size_t per = (n + nt - 1) / nt, lo = t * per, hi = lo + per < n ? lo + per : n;
    for (size_t i = lo; i < hi; ++i) {
      const bool empty = rad[i] < 0;
      if (!empty && !(mass[i] > 0)) bad[t] = 1;   // primitives.py:588
      const double px = empty ? 0.0 : pos[2 * i], py = empty ? 0.0 : pos[2 * i + 1];
      const double vx = empty ? 0.0 : vel[2 * i], vy = empty ? 0.0 : vel[2 * i + 1];
      if (f64) {
        ((double *)sp)[2 * i] = px; ((double *)sp)[2 * i + 1] = py;
        ((double *)sv)[2 * i] = vx; ((double *)sv)[2 * i + 1] = vy;
        ((double *)sr)[2 * i] = rad[i]; ((double *)sr)[2 * i + 1] = mass[i];
      } else {
        ((float *)sp)[2 * i] = (float)px; ((float *)sp)[2 * i + 1] = (float)py;
        ((float *)sv)[2 * i] = (float)vx; ((float *)sv)[2 * i + 1] = (float)vy;
        ((float *)sr)[2 * i] = (float)rad[i]; ((float *)sr)[2 * i + 1] = (float)mass[i];
      }
    }
  };
  if (nt == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto &t : th) t.join();
  }
  for (int f : bad)
    if (f) return fail(PPB_E_INVALID, "mass has to be a positive number");
  CU(cudaMemcpyAsync((char *)b->S.pos + off * 2 * esz, sp, n * 2 * esz, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync((char *)b->S.vel + off * 2 * esz, sv, n * 2 * esz, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync((char *)b->S.rm + off * 2 * esz, sr, n * 2 * esz, cudaMemcpyHostToDevice, st));
  CU(do_reset(b, world0, count, 0, st));
  CU(cudaStreamSynchronize(st));
  b->radii_known = false;
  return PPB_OK;
}

int ppb_apply_force(ppb_batch *b, int32_t world0, int32_t count, const double *force, double force_t, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  if (!force) return fail(PPB_E_INVALID, "force is null");
  if (!(force_t == force_t)) return fail(PPB_E_INVALID, "force_t is NaN");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)count * b->cfg.n_balls * 2;
  if (n == 0) return PPB_OK;
  if ((rc = stage_reserve(b, n * sizeof(double)))) return rc;
  memcpy(b->stage, force, n * sizeof(double));
  double *d_force = nullptr;
  CU(cudaMalloc((void **)&d_force, n * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(d_force, b->stage, n * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess)
    e = is64(b) ? Launch<double>::apply_force(b->S, b->cfg.n_balls, world0, count, d_force, force_t, st)
                : Launch<float>::apply_force(b->S, b->cfg.n_balls, world0, count, d_force, force_t, st);
  ++b->launches;
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_force);
  if (e != cudaSuccess) return fail(PPB_E_CUDA, "ppb_apply_force: %s", cudaGetErrorString(e));
  return PPB_OK;
}

int ppb_sample_rect_worlds(ppb_batch *b, int32_t world0, int32_t count, const ppb_sampler_params *sp, int32_t *n_failed,
                           void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  if (!sp) return fail(PPB_E_INVALID, "null params");
  if (b->cfg.n_walls < 4) return fail(PPB_E_INVALID, "the rectangular cage needs 4 wall slots");
  if (!(sp->mx_wlen >= sp->mn_wlen) || !(sp->mx_ball >= sp->mn_ball) || !(sp->mn_ball >= 1) || sp->max_tries < 1 ||
      sp->n_theta < 0 || sp->n_theta > 4 || (sp->n_theta == 0 && !(sp->arena > sp->mx_wlen + sp->w_thick)))
    return fail(PPB_E_INVALID, "inconsistent sampler parameters");
  for (int i = 0; i < sp->n_theta; ++i)
    if (!(sp->theta[i] > 0 && sp->theta[i] <= 90)) return fail(PPB_E_INVALID, "rhombus angles must be in (0, 90] degrees");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  SamplerParams P;
  P.arena = sp->arena; P.w_thick = sp->w_thick; P.mn_wlen = sp->mn_wlen; P.mx_wlen = sp->mx_wlen;
  P.mn_ball = sp->mn_ball; P.mx_ball = sp->mx_ball; P.mn_force = sp->mn_force; P.mx_force = sp->mx_force;
  P.force_t = sp->force_t; P.seed = sp->seed; P.world_index0 = sp->world_index0; P.max_tries = sp->max_tries;
  P.n_theta = sp->n_theta;
  for (int i = 0; i < 4; ++i) P.theta[i] = sp->theta[i];
  int32_t *d_fail = reinterpret_cast<int32_t *>(b->S.counters + 6);   // spare counter slot
  CU(cudaMemsetAsync(d_fail, 0, sizeof(int32_t), st));
  CU(is64(b) ? Launch<double>::sample_rect(b->S, b->cfg.n_balls, b->cfg.n_walls, world0, count, P, b->host_step, d_fail, st)
             : Launch<float>::sample_rect(b->S, b->cfg.n_balls, b->cfg.n_walls, world0, count, P, b->host_step, d_fail, st));
  CU(is64(b) ? Launch<double>::wall_cache(b->S, b->L, b->RL, world0, count, st)
             : Launch<float>::wall_cache(b->S, b->L, b->RL, world0, count, st));
  b->launches += 2;
  int32_t nf = 0;
  CU(cudaMemcpyAsync(&nf, d_fail, sizeof nf, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (n_failed) *n_failed = nf;
  b->radii_known = false;
  return PPB_OK;
}

int ppb_update_balls(ppb_batch *b, int32_t world0, int32_t count, const double *pos, const double *vel,
                     void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  if (pos && (rc = upload(b, b->S.pos, off * 2, pos, n * 2, st))) return rc;
  if (vel && (rc = upload(b, b->S.vel, off * 2, vel, n * 2, st))) return rc;
  return PPB_OK;
}

int ppb_requeue(ppb_batch *b, int32_t world0, int32_t count, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  CU(do_reset(b, world0, count, 1, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  b->radii_known = false;   // radii written through ppb_device_ptr(3) reach the renderer's records at its next call
  return PPB_OK;
}

int ppb_get_balls(ppb_batch *b, int32_t world0, int32_t count, double *pos, double *vel, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  if (pos && (rc = download(b, b->S.pos, off * 2, pos, n * 2, st))) return rc;
  if (vel && (rc = download(b, b->S.vel, off * 2, vel, n * 2, st))) return rc;
  return PPB_OK;
}

int ppb_get_radius_mass(ppb_batch *b, int32_t world0, int32_t count, double *rad, double *mass, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  std::vector<double> rm(n * 2);
  if ((rc = download(b, b->S.rm, off * 2, rm.data(), n * 2, (cudaStream_t)stream))) return rc;
  for (size_t i = 0; i < n; ++i) {
    if (rad) rad[i] = rm[2 * i];
    if (mass) mass[i] = rm[2 * i + 1];
  }
  return PPB_OK;
}

int ppb_get_walls(ppb_batch *b, int32_t world0, int32_t count, double *wall, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  if (!wall) return fail(PPB_E_INVALID, "wall is null");
  CU(cudaSetDevice(b->cfg.device));
  const size_t per = (size_t)b->cfg.n_walls * 8;
  return download(b, b->S.wall, (size_t)world0 * per, wall, (size_t)count * per, (cudaStream_t)stream);
}

int ppb_get_collision_cache(ppb_batch *b, int32_t world0, int32_t count, double *tcol, int32_t *partner,
                            void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  if (tcol && (rc = download(b, b->S.tcol, off, tcol, n, st))) return rc;
  if (partner) {
    CU(cudaMemcpyAsync(partner, b->S.partner + off, n * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PPB_OK;
}

int ppb_get_status(ppb_batch *b, int32_t world0, int32_t count, int32_t *status, int32_t *steps_done, void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<uint32_t> raw((size_t)count * 4);
  CU(cudaMemcpyAsync(raw.data(), (char *)b->S.hdr + (size_t)world0 * 16, (size_t)count * 16, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < count; ++i) {
    // Hdr<float>: tmin, step, flags, pad      Hdr<double>: tmin(2 words), step, flags
    const uint32_t step = is64(b) ? raw[4 * i + 2] : raw[4 * i + 1];
    const uint32_t flags = is64(b) ? raw[4 * i + 3] : raw[4 * i + 2];
    if (status) status[i] = (int32_t)((flags >> 8) & 0xffu);
    if (steps_done) steps_done[i] = (int32_t)step;
  }
  return PPB_OK;
}

int ppb_set_styles(ppb_batch *b, const ppb_ball_style *balls, const ppb_wall_style *walls, int32_t r_min,
                   int32_t r_max, void *stream) {
  if (!b || !balls || (!walls && b->cfg.n_walls > 0)) return fail(PPB_E_INVALID, "null argument");
  if (r_min < 0 || r_max < r_min || r_max > 4096) return fail(PPB_E_INVALID, "bad radius range [%d, %d]", r_min, r_max);
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int nB = b->cfg.n_balls, nW = b->cfg.n_walls;
  b->RL.r_min = r_min;
  b->RL.r_max = r_max;
  int max_st = 0;
  std::vector<uint32_t> fill8(nB), stroke8(nB);
  for (int i = 0; i < nB; ++i) {
    if (balls[i].s_thick < 0 || balls[i].s_thick > 1024) return fail(PPB_E_INVALID, "bad s_thick");
    b->RL.s_thick[i] = balls[i].s_thick;
    if (balls[i].s_thick > max_st) max_st = balls[i].s_thick;
    fill8[i] = h_premul_rgba8(balls[i].fill);
    stroke8[i] = h_premul_rgba8(balls[i].stroke);
  }
  for (int q = 0; q < nW; ++q) {
    b->RL.wall[q].kind = walls[q].kind;
    b->RL.wall[q].rgba8 = h_premul_rgba8(walls[q].fill);
  }
  const size_t max_sz = (size_t)2 * (r_max + max_st);
  const size_t sprite_bytes = ((max_sz * max_sz * 4 + 15) / 16) * 16;
  const size_t nr = (size_t)(r_max - r_min + 1);
  b->RL.sprite_stride = (int32_t)sprite_bytes;
  b->RL.sprite_slot_stride = (int32_t)(sprite_bytes * nr);
  const size_t need = sprite_bytes * nr * nB;
  if (need > ((size_t)1 << 31)) return fail(PPB_E_INVALID, "sprite atlas too large");
  if (need > b->atlas_bytes) {
    if (b->atlas) cudaFree(b->atlas);
    b->atlas = nullptr;
    CU(cudaMalloc((void **)&b->atlas, need));
    b->atlas_bytes = need;
  }
  if (!b->d_fill8) CU(cudaMalloc((void **)&b->d_fill8, PPB_MAX_BALLS * 4));
  if (!b->d_stroke8) CU(cudaMalloc((void **)&b->d_stroke8, PPB_MAX_BALLS * 4));
  CU(cudaMemsetAsync(b->atlas, 0, need, st));
  CU(cudaMemcpyAsync(b->d_fill8, fill8.data(), nB * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(b->d_stroke8, stroke8.data(), nB * 4, cudaMemcpyHostToDevice, st));
  CU(launch_build_sprites(b->atlas, b->RL, nB, b->d_fill8, b->d_stroke8, st));
  ++b->launches;
  // the wall kind decides the paste rectangle (primitives.py:242-245 vs 310-312)
  CU(is64(b) ? Launch<double>::wall_cache(b->S, b->L, b->RL, 0, b->cfg.n_worlds, st)
             : Launch<float>::wall_cache(b->S, b->L, b->RL, 0, b->cfg.n_worlds, st));
  ++b->launches;
  CU(cudaStreamSynchronize(st));
  b->styles_set = true;
  return PPB_OK;
}

int ppb_read_sprite(ppb_batch *b, int32_t slot, int32_t radius, uint8_t *out_host, int64_t cap_bytes, int32_t *sz_out,
                    void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (!b->styles_set) return fail(PPB_E_STATE, "ppb_set_styles must be called first");
  if (slot < 0 || slot >= b->cfg.n_balls) return fail(PPB_E_INVALID, "ball slot %d out of range", slot);
  if (radius < b->RL.r_min || radius > b->RL.r_max)
    return fail(PPB_E_INVALID, "radius %d outside the sprite range [%d, %d]", radius, b->RL.r_min, b->RL.r_max);
  const int sz = 2 * (radius + b->RL.s_thick[slot]);
  if (sz_out) *sz_out = sz;
  const size_t bytes = (size_t)sz * sz * 4;
  if (!out_host) return PPB_OK;
  if ((int64_t)bytes > cap_bytes) return fail(PPB_E_INVALID, "sprite needs %zu bytes, buffer has %lld", bytes, (long long)cap_bytes);
  CU(cudaSetDevice(b->cfg.device));
  const uint8_t *src = reinterpret_cast<const uint8_t *>(b->atlas) + (size_t)slot * b->RL.sprite_slot_stride +
                       (size_t)(radius - b->RL.r_min) * b->RL.sprite_stride;
  CU(cudaMemcpyAsync(out_host, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return PPB_OK;
}

int ppb_set_gravity(ppb_batch *b, double gx, double gy) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  b->P.gx = gx;
  b->P.gy = gy;
  b->cfg.gx = gx;
  b->cfg.gy = gy;
  return PPB_OK;
}

// The body of ppb_step_ring_async after argument checking (also the stepping loop of ppb_simulate_host).
static int run_steps(ppb_batch *b, int32_t n_steps, void *traj_dev, uint8_t *frames_dev, int32_t ring_slots,
                     int32_t render_every, cudaStream_t st) {
  const size_t frame_bytes = (size_t)b->cfg.n_worlds * b->RL.W * b->RL.H * 4;
  const size_t snap_bytes = (size_t)b->cfg.n_worlds * b->cfg.n_balls * 2 * b->esz;
  const int n_draws = frames_dev ? n_steps / render_every : 0;
  if (n_draws <= 1) {
    // no frame, or a single one: nothing to overlap with, everything on the caller's stream; the frame is drawn from
    // the live positions between two launches
    int s = 0;
    const int draw_at = n_draws == 1 ? render_every : -1;   // steps done when the frame is due
    while (s < n_steps) {
      int n = n_steps - s;
      if (n > PPB_STEPS_PER_LAUNCH) n = PPB_STEPS_PER_LAUNCH;
      if (draw_at > s && s + n > draw_at) n = draw_at - s;
      int rc = step_physics(b, s, n, traj_dev, nullptr, 0, 0, st);
      if (rc) return rc;
      s += n;
      if (s == draw_at && (rc = step_render(b, b->S.pos, frames_dev, st))) return rc;
    }
  } else {
    // Chunks of up to `draws_per_chunk` drawn steps: one advance launch on `st` leaves a position snapshot per drawn
    // step, the chunk's render launches follow on the render stream; two sets of snapshot slots alternate.  The
    // next chunk's advance launch is released when all but the last `overlap_renders` renders of this chunk are done:
    // its tail (the few warps that own a world with a long sub-step chain) then hides behind those renders, while the
    // bulk of the rendering has the SMs to itself -- beside a render launch the physics costs the step about what it
    // would cost alone, so only the tail is worth hiding.  An advance launch takes about 135 us + 28 us per step at
    // 131,072 x 8 worlds, hence long chunks: measured on B200 with a frame per step, us per step, chunks of
    // 8 / 16 / 32 draws with overlap 0: 432 / 417 / 411, overlap 2: 412 / 409 / 408, all renders overlapped:
    // 415 / 411 / 409 (PPB_CHUNK / PPB_OVERLAP in the environment select these for A/B runs).
    int rc = pipeline_reserve(b);
    if (rc) return rc;
    int per_chunk = b->draws_per_chunk;
    if (per_chunk * render_every > PPB_STEPS_PER_LAUNCH) per_chunk = PPB_STEPS_PER_LAUNCH / render_every;
    if (per_chunk < 1) per_chunk = 1;
    int n_rendered = 0, chunk_i = 0, s = 0;
    while (s < n_steps) {
      int draws = n_draws - n_rendered;
      if (draws > per_chunk) draws = per_chunk;
      // (steps after the last drawn one run without snapshots)
      int n = draws > 0 ? draws * render_every : n_steps - s;
      if (draws == 0 && n > PPB_STEPS_PER_LAUNCH) n = PPB_STEPS_PER_LAUNCH;
      const int buf = chunk_i & 1;
      if (draws > 0 && chunk_i >= 2) CU(cudaStreamWaitEvent(st, b->ev_rend[buf], 0));   // snapshot slots free again
      // how much of the previous chunk's rendering this launch may run beside (see overlap_renders)
      const bool gated = b->overlap_renders >= 0 && chunk_i >= 1;
      if (gated) CU(cudaStreamWaitEvent(st, b->ev_gate[buf ^ 1], 0));
      rc = step_physics(b, s, n, traj_dev, draws > 0 ? b->snap[buf] : nullptr, render_every - 1, render_every, st);
      if (rc) return rc;
      if (draws > 0) {
        CU(cudaEventRecord(b->ev_phys[buf], st));
        CU(cudaStreamWaitEvent(b->render_stream, b->ev_phys[buf], 0));
        const int gate_after = draws - 1 - (b->overlap_renders > 0 ? b->overlap_renders : 0);
        if (gate_after < 0) CU(cudaEventRecord(b->ev_gate[buf], b->render_stream));
        for (int d = 0; d < draws; ++d) {
          uint8_t *frame = frames_dev + (size_t)((n_rendered + d) % ring_slots) * frame_bytes;
          rc = step_render(b, (char *)b->snap[buf] + (size_t)d * snap_bytes, frame, b->render_stream);
          if (rc) return rc;
          if (d == gate_after) CU(cudaEventRecord(b->ev_gate[buf], b->render_stream));
        }
        CU(cudaEventRecord(b->ev_rend[buf], b->render_stream));
        n_rendered += draws;
        ++chunk_i;
      }
      s += n;
    }
    // the caller's stream sees every frame complete
    if (chunk_i >= 1) CU(cudaStreamWaitEvent(st, b->ev_rend[(chunk_i - 1) & 1], 0));
    if (chunk_i >= 2) CU(cudaStreamWaitEvent(st, b->ev_rend[chunk_i & 1], 0));
  }
  b->host_step += n_steps;   // the kernels take the step index by value (StepParams.step_abs)
  return PPB_OK;
}

int ppb_step_ring_async(ppb_batch *b, int32_t n_steps, void *traj_dev, uint8_t *frames_dev, int32_t ring_slots,
                        int32_t render_every, void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (n_steps < 0) return fail(PPB_E_INVALID, "n_steps must be >= 0");
  if (frames_dev) {
    if (!b->styles_set) return fail(PPB_E_STATE, "ppb_set_styles must be called before rendering");
    if (render_every < 1) return fail(PPB_E_INVALID, "render_every must be >= 1");
    if (ring_slots < 1) return fail(PPB_E_INVALID, "ring_slots must be >= 1");
    if (b->RL.W < 1 || b->RL.H < 1) return fail(PPB_E_STATE, "batch was created without a canvas");
  }
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_dev) {
    int rc = check_radii(b, st);
    if (rc) return rc;
  }
  return run_steps(b, n_steps, traj_dev, frames_dev, ring_slots, render_every, st);
}

int ppb_step_async(ppb_batch *b, int32_t n_steps, void *traj_dev, uint8_t *frames_dev, int32_t render_every,
                   void *stream) {
  const int32_t slots = (frames_dev && render_every >= 1 && n_steps / render_every > 0) ? n_steps / render_every : 1;
  return ppb_step_ring_async(b, n_steps, traj_dev, frames_dev, slots, render_every, stream);
}

int ppb_scan_async(ppb_batch *b, void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  StepParams P = b->P;
  P.step_off = 0;
  P.step_abs = b->host_step;
  P.scan_only = 1;
  P.n_steps = 1;
  P.draw_first = 0;
  P.draw_every = 0;
  P.parity = b->launch_parity;
  b->launch_parity ^= 1;
  CU(is64(b) ? Launch<double>::advance(b->S, b->L, P, nullptr, b->advance_grid, st)
             : Launch<float>::advance(b->S, b->L, P, nullptr, b->advance_grid, st));
  b->launches += 1;
  return PPB_OK;
}

int ppb_get_collision_response(ppb_batch *b, int32_t world0, int32_t count, double *nrml, double *future_vel,
                               void *stream) {
  int rc = check_range(b, world0, count);
  if (rc) return rc;
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nB = (size_t)b->cfg.n_balls, n = (size_t)count * nB, off = (size_t)world0 * nB;
  std::vector<double> aux(n * 4);
  if ((rc = download(b, b->S.aux, off * 4, aux.data(), n * 4, st))) return rc;
  for (size_t i = 0; i < n; ++i) {
    if (nrml) { nrml[2 * i] = aux[4 * i]; nrml[2 * i + 1] = aux[4 * i + 1]; }
    if (future_vel) { future_vel[2 * i] = aux[4 * i + 2]; future_vel[2 * i + 1] = aux[4 * i + 3]; }
  }
  return PPB_OK;
}

int ppb_render_async(ppb_batch *b, uint8_t *frames_dev, void *stream) {
  if (!b || !frames_dev) return fail(PPB_E_INVALID, "null argument");
  if (!b->styles_set) return fail(PPB_E_STATE, "ppb_set_styles must be called before rendering");
  if (b->RL.W < 1 || b->RL.H < 1) return fail(PPB_E_STATE, "batch was created without a canvas");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  {
    int rc = check_radii(b, st);
    if (rc) return rc;
  }
  {
    Prof p(b, st, 2);
    CU(is64(b) ? Launch<double>::render(b->S, b->L, b->RL, b->atlas, frames_dev, st)
               : Launch<float>::render(b->S, b->L, b->RL, b->atlas, frames_dev, st));
    ++b->launches;
  }
  return PPB_OK;
}

static int simulate_host_impl(ppb_batch *b, int32_t n_steps, float *traj_host, uint8_t *frames_host,
                              int32_t render_every, void *stream, bool wait) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (n_steps < 0) return fail(PPB_E_INVALID, "n_steps must be >= 0");
  if (frames_host) {
    if (!b->styles_set) return fail(PPB_E_STATE, "ppb_set_styles must be called before rendering");
    if (render_every < 1) return fail(PPB_E_INVALID, "render_every must be >= 1");
    if (b->RL.W < 1 || b->RL.H < 1) return fail(PPB_E_STATE, "batch was created without a canvas");
  } else {
    render_every = 1;
  }
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  if (frames_host) {
    int rc = check_radii(b, st);
    if (rc) return rc;
  }
  if (frames_host && b->pack_frames && !b->pack) {
    b->pack = packpipe_create(b->cfg.device);
    if (!b->pack) return fail(PPB_E_CUDA, "could not set up the packed frame download");
  }
  if (b->pack) b->pack_mark[b->sim_calls & 1u] = packpipe_submitted(b->pack);
  const size_t nW = (size_t)b->cfg.n_worlds, nB = (size_t)b->cfg.n_balls;
  const size_t frame_bytes = nW * b->RL.W * b->RL.H * 4;
  const size_t traj_elems = nW * nB * 2;  // per step
  // chunk = a whole number of render periods, bounded so that the two device staging buffers
  // stay below ~1 GiB each (but at least one period)
  int chunk = render_every;
  {
    const size_t per_step = (traj_host ? traj_elems * (b->esz + 4) : 0);
    const size_t per_period = per_step * render_every + (frames_host ? frame_bytes : 0);
    size_t periods = per_period ? ((size_t)1 << 30) / per_period : 64;
    if (periods < 1) periods = 1;
    if (periods > 64) periods = 64;
    chunk = (int)(periods * render_every);
  }
  const size_t need_traj = traj_host ? traj_elems * chunk * b->esz : 0;
  const size_t need_frames = frames_host ? frame_bytes * (chunk / render_every) : 0;
  if (chunk != b->scr_chunk || need_traj > b->scr_traj_bytes || need_frames > b->scr_frames_bytes) {
    CU(cudaDeviceSynchronize());
    if (b->pack) {   // the frame downloads in flight read the scratch that is about to go
      std::string perr;
      if (packpipe_wait_finished(b->pack, packpipe_submitted(b->pack), &perr))
        return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
    }
    free_scratch(b);
    for (int i = 0; i < 2; ++i) {
      if (need_traj) {
        CU(cudaMalloc(&b->scr_traj[i], need_traj));
        if (is64(b)) CU(cudaMalloc((void **)&b->scr_traj32[i], need_traj / 2));
      }
      if (need_frames) CU(cudaMalloc((void **)&b->scr_frames[i], need_frames));
      CU(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b->ev_copied[i], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&b->ev_call[i], cudaEventDisableTiming));
    }
    CU(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    b->scr_traj_bytes = need_traj;
    b->scr_frames_bytes = need_frames;
    b->scr_chunk = chunk;
  }
  int done = 0;
  while (done < n_steps) {
    const int cur = (n_steps - done < chunk) ? (n_steps - done) : chunk;
    const int buf = (int)(b->sim_it & 1u);
    if (b->sim_it >= 2) CU(cudaStreamWaitEvent(st, b->ev_copied[buf], 0));  // staging buffer free again
    if (frames_host && b->pack) {   // ... and the frames the packed download of this slot's previous chunk reads
      std::string perr;
      if (packpipe_wait_slot(b->pack, buf, &perr)) return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
    }
    {
      // chunk boundaries are multiples of render_every, so the frames of this chunk fill scr_frames[buf] in order;
      // run_steps pipelines the renders with the physics of the following steps and advances host_step
      const int nf = frames_host ? cur / render_every : 0;
      int rc = run_steps(b, cur, traj_host ? b->scr_traj[buf] : nullptr, nf > 0 ? b->scr_frames[buf] : nullptr,
                         nf > 0 ? nf : 1, render_every, st);
      if (rc) return rc;
    }
    const float *src32 = (const float *)b->scr_traj[buf];
    if (traj_host && is64(b)) {
      CU(Launch<double>::traj_to_f32(b->scr_traj[buf], b->scr_traj32[buf], (long long)traj_elems * cur, st));
      ++b->launches;
      src32 = b->scr_traj32[buf];
    }
    CU(cudaEventRecord(b->ev_done[buf], st));
    CU(cudaStreamWaitEvent(b->copy_stream, b->ev_done[buf], 0));
    if (traj_host)
      CU(cudaMemcpyAsync(traj_host + (size_t)done * traj_elems, src32, traj_elems * cur * 4, cudaMemcpyDeviceToHost,
                         b->copy_stream));
    if (frames_host) {
      const int f0 = done / render_every, nf = (done + cur) / render_every - f0;
      if (nf > 0 && b->pack) {
        std::string perr;
        const int nl = packpipe_submit(b->pack, buf, b->scr_frames[buf], (long long)nf * (long long)nW, b->RL.W * b->RL.H,
                                       frames_host + (size_t)f0 * frame_bytes, st, &perr);
        if (nl < 0) return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
        b->launches += nl;
      } else if (nf > 0) {
        CU(cudaMemcpyAsync(frames_host + (size_t)f0 * frame_bytes, b->scr_frames[buf], frame_bytes * nf,
                           cudaMemcpyDeviceToHost, b->copy_stream));
      }
    }
    CU(cudaEventRecord(b->ev_copied[buf], b->copy_stream));
    done += cur;
    ++b->sim_it;
  }
  CU(cudaEventRecord(b->ev_call[b->sim_calls & 1u], b->copy_stream));
  ++b->sim_calls;
  if (wait) {
    CU(cudaStreamSynchronize(st));
    CU(cudaStreamSynchronize(b->copy_stream));
    if (b->pack) {
      std::string perr;
      if (packpipe_wait_finished(b->pack, packpipe_submitted(b->pack), &perr))
        return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
    }
  }
  return PPB_OK;
}

int ppb_simulate_host(ppb_batch *b, int32_t n_steps, float *traj_host, uint8_t *frames_host, int32_t render_every,
                      void *stream) {
  return simulate_host_impl(b, n_steps, traj_host, frames_host, render_every, stream, true);
}

int ppb_simulate_host_async(ppb_batch *b, int32_t n_steps, float *traj_host, uint8_t *frames_host,
                            int32_t render_every, void *stream) {
  return simulate_host_impl(b, n_steps, traj_host, frames_host, render_every, stream, false);
}

int ppb_download_frames(ppb_batch *b, const uint8_t *frames_dev, int64_t n_frames, uint8_t *frames_host, void *stream) {
  if (!b || !frames_dev || !frames_host) return fail(PPB_E_INVALID, "null argument");
  if (n_frames < 0) return fail(PPB_E_INVALID, "n_frames must be >= 0");
  if (b->RL.W < 1 || b->RL.H < 1) return fail(PPB_E_STATE, "batch was created without a canvas");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t frame_bytes = (size_t)b->RL.W * b->RL.H * 4;
  if (!b->pack_frames) {
    CU(cudaMemcpyAsync(frames_host, frames_dev, frame_bytes * (size_t)n_frames, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return PPB_OK;
  }
  if (!b->pack && !(b->pack = packpipe_create(b->cfg.device))) return fail(PPB_E_CUDA, "could not set up the packed frame download");
  std::string perr;
  // pieces of at most 1 GiB of raw frames (the token buffers are sized by the piece), alternating job slots
  const int64_t per = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / frame_bytes));
  int slot = 0;
  for (int64_t f0 = 0; f0 < n_frames; f0 += per, slot ^= 1) {
    const int64_t nf = std::min<int64_t>(per, n_frames - f0);
    const int nl = packpipe_submit(b->pack, slot, frames_dev + (size_t)f0 * frame_bytes, nf, b->RL.W * b->RL.H,
                                   frames_host + (size_t)f0 * frame_bytes, st, &perr);
    if (nl < 0) return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
    b->launches += nl;
  }
  if (packpipe_wait_finished(b->pack, packpipe_submitted(b->pack), &perr))
    return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
  return PPB_OK;
}

int ppb_host_sync(ppb_batch *b, int32_t keep_in_flight, void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (keep_in_flight < 0 || keep_in_flight > 1) return fail(PPB_E_INVALID, "keep_in_flight must be 0 or 1");
  CU(cudaSetDevice(b->cfg.device));
  if (keep_in_flight == 1) {
    // everything but the most recent call: the copy stream is first-in first-out, so the end of the call before
    // the last implies all earlier ones (and the kernels they copied from)
    if (b->sim_calls >= 2 && b->ev_call[b->sim_calls & 1u]) CU(cudaEventSynchronize(b->ev_call[b->sim_calls & 1u]));
    if (b->pack && b->sim_calls >= 1) {   // the frames of every call before the last one are in the caller's arrays
      std::string perr;
      if (packpipe_wait_finished(b->pack, b->pack_mark[(b->sim_calls - 1) & 1u], &perr))
        return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
    }
    return PPB_OK;
  }
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  if (b->copy_stream) CU(cudaStreamSynchronize(b->copy_stream));
  if (b->pack) {
    std::string perr;
    if (packpipe_wait_finished(b->pack, packpipe_submitted(b->pack), &perr))
      return fail(PPB_E_CUDA, "frame download: %s", perr.c_str());
  }
  return PPB_OK;
}

int ppb_crop_async(ppb_batch *b, const uint8_t *frames_dev, const void *traj_dev, int32_t n_frames, int32_t crop_sz,
                   int32_t mode, uint8_t *crops_dev, float *pos_dev, void *stream) {
  if (!b || !frames_dev || !traj_dev || !crops_dev) return fail(PPB_E_INVALID, "null argument");
  if (n_frames < 0 || crop_sz < 1 || crop_sz > 4096 || mode < 0 || mode > 1) return fail(PPB_E_INVALID, "bad n_frames / crop_sz / mode");
  if (b->RL.W < 1 || b->RL.H < 1) return fail(PPB_E_STATE, "batch was created without a canvas");
  CU(cudaSetDevice(b->cfg.device));
  CU(is64(b) ? Launch<double>::crop(frames_dev, traj_dev, n_frames, b->cfg.n_worlds, b->cfg.n_balls, b->RL.W, b->RL.H,
                                    crop_sz, mode, crops_dev, pos_dev, (cudaStream_t)stream)
             : Launch<float>::crop(frames_dev, traj_dev, n_frames, b->cfg.n_worlds, b->cfg.n_balls, b->RL.W, b->RL.H,
                                   crop_sz, mode, crops_dev, pos_dev, (cudaStream_t)stream));
  ++b->launches;
  return PPB_OK;
}

int ppb_resize_crops_async(ppb_batch *b, const uint8_t *crops_dev, const void *traj_dev, int32_t n_frames, int32_t crop_sz,
                           int32_t proc_sz, uint8_t *out_dev, float *pos_dev, float *vel_dev, void *stream) {
  if (!b || !crops_dev || !out_dev) return fail(PPB_E_INVALID, "null argument");
  if (vel_dev && !traj_dev) return fail(PPB_E_INVALID, "displacements need the trajectory");
  if (n_frames < 0 || crop_sz < 1 || proc_sz < 1 || crop_sz > 255 || proc_sz > 1024) return fail(PPB_E_INVALID, "bad n_frames / crop_sz / proc_sz");
  CU(cudaSetDevice(b->cfg.device));
  const long long per = (long long)b->cfg.n_worlds * b->cfg.n_balls;
  cudaError_t e = is64(b) ? Launch<double>::resize_crops(crops_dev, traj_dev, n_frames, per, crop_sz, proc_sz, out_dev, pos_dev,
                                                         vel_dev, (cudaStream_t)stream)
                          : Launch<float>::resize_crops(crops_dev, traj_dev, n_frames, per, crop_sz, proc_sz, out_dev, pos_dev,
                                                        vel_dev, (cudaStream_t)stream);
  if (e == cudaErrorInvalidValue) return fail(PPB_E_INVALID, "crop_sz %d -> proc_sz %d is outside what the resize kernel holds in shared memory", crop_sz, proc_sz);
  CU(e);
  ++b->launches;
  return PPB_OK;
}

int ppb_horizon_async(ppb_batch *b, const void *traj_dev, int32_t n_rows, int32_t lookahead, float *out_dev, void *stream) {
  if (!b || !traj_dev || !out_dev) return fail(PPB_E_INVALID, "null argument");
  if (lookahead < 1 || n_rows <= lookahead) return fail(PPB_E_INVALID, "need n_rows > lookahead >= 1");
  CU(cudaSetDevice(b->cfg.device));
  const long long per = (long long)b->cfg.n_worlds * b->cfg.n_balls;
  CU(is64(b) ? Launch<double>::horizon(traj_dev, per, n_rows, lookahead, out_dev, (cudaStream_t)stream)
             : Launch<float>::horizon(traj_dev, per, n_rows, lookahead, out_dev, (cudaStream_t)stream));
  ++b->launches;
  return PPB_OK;
}

int ppb_collision_flags_async(const float *traj_dev, int64_t n_items, int32_t n_steps, float thresh, uint8_t *flags_dev,
                              void *stream) {
  if (!traj_dev || !flags_dev) return fail(PPB_E_INVALID, "null argument");
  if (n_items < 0 || n_steps < 3) return fail(PPB_E_INVALID, "need n_steps >= 3");
  CU(launch_collision_flags(traj_dev, n_items, n_steps, thresh, flags_dev, (cudaStream_t)stream));
  return PPB_OK;
}

int ppb_read_events(ppb_batch *b, ppb_event *out, int64_t cap, int64_t *n_total, void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  CU(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long n = 0;
  CU(cudaMemcpyAsync(&n, b->S.counters, sizeof n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (n_total) *n_total = (int64_t)n;
  int64_t m = (int64_t)n;
  if (m > b->cfg.event_capacity) m = b->cfg.event_capacity;
  if (m > cap) m = cap;
  if (out && m > 0) {
    CU(cudaMemcpyAsync(out, b->S.events, (size_t)m * sizeof(ppb_event), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return PPB_OK;
}

int ppb_clear_events(ppb_batch *b, void *stream) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  CU(cudaSetDevice(b->cfg.device));
  CU(cudaMemsetAsync(b->S.counters, 0, sizeof(unsigned long long), (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  return PPB_OK;
}

int ppb_get_counters(ppb_batch *b, int64_t out[8], void *stream) {
  if (!b || !out) return fail(PPB_E_INVALID, "null argument");
  CU(cudaSetDevice(b->cfg.device));
  unsigned long long c[4];
  CU(cudaMemcpyAsync(c, b->S.counters, sizeof c, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(cudaStreamSynchronize((cudaStream_t)stream));
  out[0] = b->launches;
  out[1] = (int64_t)c[1];
  out[2] = (int64_t)c[2];
  out[3] = (int64_t)c[0];
  out[4] = (int64_t)c[3];
  out[5] = out[6] = out[7] = 0;
  if (b->pack) {   // frame download: bytes that crossed PCIe packed / raw, time spent rebuilding pixels on the host
    long long ps[6];
    packpipe_stats(b->pack, ps);
    out[5] = ps[2];
    out[6] = ps[3];
    out[7] = ps[5];
  }
  return PPB_OK;
}

int ppb_device_ptr(ppb_batch *b, int32_t which, void **ptr, int64_t *nbytes) {
  if (!b || !ptr) return fail(PPB_E_INVALID, "null argument");
  const size_t nb = (size_t)b->cfg.n_worlds * b->cfg.n_balls;
  void *p = nullptr;
  size_t n = 0;
  switch (which) {
    case 0: p = b->S.pos; n = nb * 2 * b->esz; break;
    case 1: p = b->S.tcol; n = nb * b->esz; break;
    case 2: p = b->S.wall; n = (size_t)b->cfg.n_worlds * b->cfg.n_walls * 8 * b->esz; break;
    case 3: p = b->S.rm; n = nb * 2 * b->esz; break;
    case 4: p = b->S.vel; n = nb * 2 * b->esz; break;
    default: return fail(PPB_E_INVALID, "unknown buffer id %d", which);
  }
  *ptr = p;
  if (nbytes) *nbytes = (int64_t)n;
  return PPB_OK;
}

int ppb_set_profiling(ppb_batch *b, int32_t on) {
  if (!b) return fail(PPB_E_INVALID, "null batch");
  if (on < 0 || on > 3) return fail(PPB_E_INVALID, "profiling mode must be 0..3");
  static const int mask[4] = {0, 7, 4, 3};
  b->profiling = mask[on];
  b->ev_used = 0;
  return PPB_OK;
}

int ppb_get_kernel_times(ppb_batch *b, float ms[3], int32_t launches[3]) {
  if (!b || !ms || !launches) return fail(PPB_E_INVALID, "null argument");
  for (int i = 0; i < 3; ++i) {
    ms[i] = 0.f;
    launches[i] = 0;
  }
  for (size_t i = 0; i + 1 < b->ev_used; i += 2) {
    float t = 0.f;
    CU(cudaEventSynchronize(b->ev_pool[i + 1]));
    CU(cudaEventElapsedTime(&t, b->ev_pool[i], b->ev_pool[i + 1]));
    const int kind = b->ev_kind[i / 2];
    ms[kind] += t;
    launches[kind] += 1;
  }
  b->ev_used = 0;
  return PPB_OK;
}

}  // extern "C"
