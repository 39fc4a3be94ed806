// ppb_launch_impl.cuh -- body of Launch<T>, included by ppb_f32.cu / ppb_f64.cu
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "ppb_kernels.cuh"
#include "ppb_launch.h"
#include "ppb_post.cuh"
#include "ppb_render.cuh"

namespace ppb {

// Launch with programmatic stream serialisation (PDL): the grid may be scheduled while the previous kernel
// of the stream drains; the kernel itself waits (griddepcontrol.wait) before it touches global memory.
// PPB_NO_PDL=1 in the environment falls back to plain launches (for A/B timing).
static inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("PPB_NO_PDL");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

// Tensor map over a frame buffer seen as `n_lines` rows of 128 bytes (32 RGBA pixels), box = `box_lines` rows, stored
// with the copy engine's 128-byte swizzle (render_tile_kernel, mode PPB_RM_SWZ).  The encoder is a driver function;
// the library links the static runtime only, so it is looked up once through the runtime.
static inline cudaError_t encode_frame_lines_map(CUtensorMap *out, void *frames, unsigned long long n_lines, unsigned box_lines) {
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return e;
    if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
    encode = reinterpret_cast<encode_fn>(fn);
  }
  const cuuint64_t dims[2] = {32, n_lines};
  const cuuint64_t strides[1] = {128};            // bytes between rows
  const cuuint32_t box[2] = {32, box_lines};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, frames, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename T>
size_t Launch<T>::hdr_size() { return sizeof(Hdr<T>); }

template <typename T>
int Launch<T>::advance_grid(int n_worlds, int n_balls, int sm_count) {
  int occ = 0;
  if (n_balls <= 8) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, advance_kernel<T, 8, 1>, PPB_COLLIDE_WARPS * 32, 0);
  else if (n_balls <= 16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, advance_kernel<T, 16, 2>, PPB_COLLIDE_WARPS * 32, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, advance_kernel<T, 32, 4>, PPB_COLLIDE_WARPS * 32, 0);
  if (occ < 1) occ = 1;
  if (getenv("PPB_ADV_CARVEOUT")) {   // developer switch (A/B timing): same L1 / shared-memory split as the render kernel
    cudaFuncSetAttribute(advance_kernel<T, 1, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(advance_kernel<T, 2, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(advance_kernel<T, 4, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(advance_kernel<T, 8, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(advance_kernel<T, 16, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(advance_kernel<T, 32, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  if (const char *e = getenv("PPB_ADV_OCC")) { const int v = atoi(e); if (v >= 1 && v < occ) occ = v; }   // developer switch (A/B timing)
  const int tw = n_balls <= 8 ? 1 : (n_balls <= 16 ? 2 : 4);
  const int teams = PPB_COLLIDE_WARPS / tw;
  const int need = (n_worlds + teams - 1) / teams;   // at least one world per team
  int grid = sm_count * occ;
  if (grid > need) grid = need;
  return grid < 1 ? 1 : grid;
}

template <typename T>
cudaError_t Launch<T>::advance(const StatePtrs &S, const Layout &L, const StepParams &P, void *traj, int grid,
                               cudaStream_t st) {
  const int nb = L.n_balls;
  const dim3 blk(PPB_COLLIDE_WARPS * 32);
  // lanes per world: one warp up to 8 balls, a team of 2 warps for 16, of 4 warps for 32
  if (nb <= 1) return launch_pdl(advance_kernel<T, 1, 1>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
  else if (nb <= 2) return launch_pdl(advance_kernel<T, 2, 1>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
  else if (nb <= 4) return launch_pdl(advance_kernel<T, 4, 1>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
  else if (nb <= 8) return launch_pdl(advance_kernel<T, 8, 1>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
  else if (nb <= 16) return launch_pdl(advance_kernel<T, 16, 2>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
  else return launch_pdl(advance_kernel<T, 32, 4>, dim3(grid), blk, 0, st, S, L, P, (T *)traj);
}

template <typename T>
cudaError_t Launch<T>::sample_rect(const StatePtrs &S, int n_balls, int n_walls, int world0, int count,
                                   const SamplerParams &P, int32_t step_value, int32_t *fail_count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  sample_rect_worlds_kernel<T><<<(count + 127) / 128, 128, 0, st>>>(S, n_balls, n_walls, world0, count, P, step_value, fail_count);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::crop(const uint8_t *frames, const void *traj, long long n_frames, int n_worlds, int n_balls, int W,
                            int H, int c, int mode, uint8_t *crops, float *pos_out, cudaStream_t st) {
  const long long items = n_frames * n_worlds * n_balls;
  if (items <= 0) return cudaSuccess;
  if (items > 2147483647LL) return cudaErrorInvalidValue;
  const int threads = c * c >= 256 ? 256 : (c * c >= 128 ? 128 : 64);
  crop_kernel<T><<<(unsigned)items, threads, 0, st>>>(frames, (const T *)traj, n_worlds, n_balls, W, H, c, mode, crops, pos_out);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::resize_crops(const uint8_t *crops, const void *traj, long long n_frames, long long items_per_frame,
                                    int c, int p, uint8_t *out, float *pos, float *vel_out, cudaStream_t st) {
  const long long items = n_frames * items_per_frame;
  if (items <= 0) return cudaSuccess;
  if (items > 2147483647LL) return cudaErrorInvalidValue;
  const double scale = (double)c / (double)p;
  const int ksize = (int)ceil(scale < 1.0 ? 1.0 : scale) * 2 + 1;
  if (ksize > 32) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(2 * p + p * ksize) * sizeof(int) + (size_t)3 * c * c + (size_t)3 * c * p;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute(resize_crops_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  resize_crops_kernel<T><<<(unsigned)items, 128, smem, st>>>(crops, (const T *)traj, items_per_frame, c, p, ksize, out, pos, vel_out);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::horizon(const void *traj, long long n_items, int n_rows, int lookahead, float *out, cudaStream_t st) {
  const long long total = (long long)(n_rows - lookahead) * n_items * lookahead;
  if (total <= 0) return cudaSuccess;
  horizon_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T *)traj, n_items, n_rows, lookahead, out);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::wall_cache(const StatePtrs &S, const Layout &L, const RenderLayout &RL, int world0, int count,
                                  cudaStream_t st) {
  const long long n = (long long)count * L.n_walls;
  if (n <= 0) return cudaSuccess;
  wall_cache_kernel<T><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(S, L.n_walls, L.n_balls, RL, world0, count);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::render(const StatePtrs &S, const Layout &L, const RenderLayout &RL, const uint32_t *atlas,
                              uint8_t *frames, cudaStream_t st) {
  const int tiles = ((RL.W + PPB_TILE - 1) / PPB_TILE) * ((RL.H + PPB_TILE - 1) / PPB_TILE);
  long long jobs = (long long)S.n_worlds * tiles;
  if (jobs == 0) return cudaSuccess;
  // fast path: every wall precedes every ball in the insertion order and rows are 16-byte multiples
  bool walls_first = true;
  for (int k = 0; k < L.n_obj; ++k)
    if (L.order[k] < L.n_walls && k >= L.n_walls) walls_first = false;
  if (walls_first && (reinterpret_cast<uintptr_t>(frames) & 3u) == 0) {
    // canvases other than 64 pixels wide: tiles are bands of whole frame rows
    static const bool no_band = getenv("PPB_NO_BAND") != nullptr;   // developer switch: square tiles everywhere (A/B timing)
    const bool band = RL.W != PPB_TILE && ((RL.W + 3) & ~3) <= PPB_TILE * PPB_TILE && !no_band;
    if (band) {
      const int rh = (PPB_TILE * PPB_TILE) / ((RL.W + 3) & ~3);
      jobs = (long long)S.n_worlds * ((RL.H + rh - 1) / rh);
    }
    const int warps = render_warps(L.n_walls, L.n_balls);
    const int smem = (int)(warps * render_warp_smem_bytes(L.n_walls, L.n_balls));
    // the 64-pixel canvas: swizzled tile, tensor-map store (frames as an array of 128-byte lines, box = one frame)
    static const bool no_swz = getenv("PPB_NO_SWIZZLE") != nullptr;   // developer switch (A/B timing)
    const bool swz = !band && !no_swz && RL.W == PPB_TILE && RL.H <= PPB_TILE &&
                     (reinterpret_cast<uintptr_t>(frames) & 15u) == 0 && (long long)S.n_worlds * 2 * RL.H <= 2147483647LL;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (swz) {
      cudaError_t e = encode_frame_lines_map(&tmap, frames, (unsigned long long)S.n_worlds * 2ull * (unsigned)RL.H, 2u * (unsigned)RL.H);
      if (e != cudaSuccess) return e;
    }
    const int smem_max = 227 * 1024;
    // a small sprite atlas rides along in shared memory (the 64-pixel canvas only)
    static const bool no_satl = getenv("PPB_NO_SMEM_ATLAS") != nullptr;   // developer switch (A/B timing)
    const int atlas_bytes = L.n_balls * RL.sprite_slot_stride;
    const bool satl = swz && !no_satl && atlas_bytes > 0 && atlas_bytes <= 8192 && (atlas_bytes & 3) == 0 && smem + atlas_bytes <= smem_max;
    // per device (and per instantiation T): SM count, opt-in to the large dynamic shared memory
    static int sm_counts[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (!sm_counts[dev]) {
      int n = 0;
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      cudaError_t e = cudaFuncSetAttribute(render_tile_kernel<T, PPB_RM_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(render_tile_kernel<T, PPB_RM_BAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(render_tile_kernel<T, PPB_RM_SWZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(render_tile_kernel<T, PPB_RM_SWZ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
      if (e != cudaSuccess) return e;
      sm_counts[dev] = n > 0 ? n : 1;
    }
    const int sm_count = sm_counts[dev];
    // ticket counters of the dynamic tile hand-out: a ring of {next ticket, warps done} pairs per device, so that render
    // launches in flight on different streams never share one; a launch leaves its pair zeroed
    constexpr int kCtrRing = 256;
    static uint32_t *ctr_ring[64] = {nullptr};
    static std::atomic<unsigned> ctr_next[64];
    static std::mutex ctr_mu;
    {
      std::lock_guard<std::mutex> lk(ctr_mu);   // (host threads with batches of their own may render at the same time)
      if (!ctr_ring[dev]) {
        uint32_t *ring = nullptr;
        cudaError_t e = cudaMalloc((void **)&ring, kCtrRing * 2 * sizeof(uint32_t));
        if (e == cudaSuccess) e = cudaMemset(ring, 0, kCtrRing * 2 * sizeof(uint32_t));
        if (e != cudaSuccess) return e;
        ctr_ring[dev] = ring;
      }
    }
    uint32_t *tile_ctr = ctr_ring[dev] + 2 * (ctr_next[dev].fetch_add(1u) % kCtrRing);
    static const bool static_split = getenv("PPB_RENDER_STATIC") != nullptr;   // developer switch (A/B timing)
    if (static_split) tile_ctr = nullptr;
    // tiles per ticket (see the kernel; 4 - 8 measured alike, 346 - 348 us per launch of 131,072 frames)
    static const uint32_t chunk = getenv("PPB_RENDER_CHUNK") ? (uint32_t)atoi(getenv("PPB_RENDER_CHUNK")) : 6u;
    // persistent: one CTA of `warps` warps per SM, every warp walks the tile list
    long long blocks = (jobs + warps - 1) / warps;
    if (blocks > sm_count) blocks = sm_count;
    if (band) render_tile_kernel<T, PPB_RM_BAND><<<(unsigned)blocks, warps * 32, smem, st>>>(S, L, RL, atlas, frames, tmap, tile_ctr, chunk);
    else if (swz && satl) render_tile_kernel<T, PPB_RM_SWZ, true><<<(unsigned)blocks, warps * 32, smem + atlas_bytes, st>>>(S, L, RL, atlas, frames, tmap, tile_ctr, chunk);
    else if (swz) render_tile_kernel<T, PPB_RM_SWZ><<<(unsigned)blocks, warps * 32, smem, st>>>(S, L, RL, atlas, frames, tmap, tile_ctr, chunk);
    else render_tile_kernel<T, PPB_RM_TILE><<<(unsigned)blocks, warps * 32, smem, st>>>(S, L, RL, atlas, frames, tmap, tile_ctr, chunk);
  } else {
    if (jobs > 2147483647LL) return cudaErrorInvalidValue;
    render_general_kernel<T><<<(unsigned)jobs, 256, 0, st>>>(S, L, RL, atlas, frames);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::reset_cache(const StatePtrs &S, int n_balls, int world0, int count, int keep_cache,
                                   int32_t step_value, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  reset_cache_kernel<T><<<(count + 127) / 128, 128, 0, st>>>(S, n_balls, world0, count, keep_cache, step_value);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::apply_force(const StatePtrs &S, int n_balls, int world0, int count, const double *force,
                                   double force_t, cudaStream_t st) {
  const long long n = (long long)count * n_balls;
  if (n <= 0) return cudaSuccess;
  apply_force_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, n_balls, world0, count, force, force_t);
  return cudaGetLastError();
}

template <typename T>
cudaError_t Launch<T>::radius_range(const StatePtrs &S, int n_walls, int n_balls, int32_t *out, cudaStream_t st) {
  const long long n = (long long)S.n_worlds * n_balls;
  if (n <= 0) return cudaSuccess;
  radius_range_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, n_walls, n_balls, out);
  return cudaGetLastError();
}

template <typename T>
__global__ void traj_to_f32_kernel(const T *src, float *dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (float)src[i];
}

template <typename T>
cudaError_t Launch<T>::traj_to_f32(const void *src, float *dst, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  traj_to_f32_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T *)src, dst, n);
  return cudaGetLastError();
}

}  // namespace ppb
