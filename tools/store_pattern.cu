// store_pattern.cu -- does the ORDER in which 16 KiB tiles reach memory matter?  (The band renderer writes 14,400-byte
// tiles of 1200-pixel canvases at 7.2 TB/s, the 16 KiB skeleton of store_ceiling.cu stops at 6.4 TB/s.)
// Variants of the bulk-store skeleton over the same bytes: tile size, pieces per tile issued in rotated order, tile->warp maps.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/store_pattern tools/store_pattern.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

// map: 0 = round robin (tile = cta*warps + warp + k*stride), 1 = each warp owns a contiguous range of tiles,
//      2 = each CTA owns a contiguous range, its warps round-robin inside it
// pieces: a tile leaves as `pieces` bulk copies of tile_bytes/pieces, the first one chosen by (tile + warp) % pieces if rot
__global__ void bulk_kernel(uint8_t *dst, long long n_tiles, int warps, int tile_bytes, int pieces, int rot, int map, int depth) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *tile = smem + (size_t)warp * 16384;
  for (int i = lane * 16; i < 16384; i += 32 * 16) *reinterpret_cast<uint4 *>(tile + i) = make_uint4(0xff0000ffu, ~0u, ~0u, ~0u);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
  const long long n_warps = (long long)gridDim.x * warps;
  const long long gw = (long long)blockIdx.x * warps + warp;
  long long t0, t1, ts;
  if (map == 0) { t0 = gw; t1 = n_tiles; ts = n_warps; }
  else if (map == 1) { const long long per = (n_tiles + n_warps - 1) / n_warps; t0 = gw * per; t1 = min(n_tiles, t0 + per); ts = 1; }
  else { const long long per = (n_tiles + gridDim.x - 1) / gridDim.x; t0 = blockIdx.x * per + warp; t1 = min(n_tiles, (long long)(blockIdx.x + 1) * per); ts = warps; }
  const int pb = tile_bytes / pieces;
  for (long long t = t0; t < t1; t += ts) {
    if (lane == 0) {
      const int first = rot ? (int)((t * 5 + warp * 3 + blockIdx.x) % pieces) : 0;
      for (int p = 0; p < pieces; ++p) {
        const int q = (first + p) % pieces;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * tile_bytes + (long long)q * pb), "r"(saddr + q * pb), "r"(pb) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (depth == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // (not legal for a renderer with one buffer: ceiling only)
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the renderer's hand-out: a warp's first tile from its grid index, then tickets; a ticket = chunk tiles that lie A/chunk apart
__global__ void ticket_kernel(uint8_t *dst, long long n_tiles, int warps, int tile_bytes, unsigned chunk, unsigned *ctr) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *tile = smem + (size_t)warp * 16384;
  for (int i = lane * 16; i < 16384; i += 32 * 16) *reinterpret_cast<uint4 *>(tile + i) = make_uint4(0xff0000ffu, ~0u, ~0u, ~0u);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
  const long long n_static = (long long)gridDim.x * warps;
  const long long n_dyn = n_tiles - n_static;
  const long long A_per = (n_dyn - 3 * n_static) / chunk, A = A_per * chunk;
  long long job = (long long)blockIdx.x * warps + warp;
  unsigned left = 0;
  long long cj = 0;
  while (job < n_tiles) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + job * tile_bytes), "r"(saddr), "r"(tile_bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (left == 0) {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(ctr, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t < A_per) { cj = n_static + t; left = chunk; } else { cj = n_static + A + (t - A_per); left = 1; }
    }
    --left;
    job = cj;
    cj += A_per;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
  const long long bytes = argc > 1 ? atoll(argv[1]) : (2LL << 30);
  uint8_t *buf;
  CK(cudaMalloc(&buf, bytes + (1 << 20)));
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 16384));
  const int reps = 20;
  struct V { int tile_bytes, pieces, rot, map, depth, warps; };
  const V vs[] = {
    {16384, 1, 0, 0, 0, 12}, {14400, 1, 0, 0, 0, 12}, {15360, 1, 0, 0, 0, 12}, {16384 - 256, 1, 0, 0, 0, 12}, {12288, 1, 0, 0, 0, 12}, {8192, 1, 0, 0, 0, 12},
    {16384, 4, 0, 0, 0, 12}, {16384, 4, 1, 0, 0, 12}, {16384, 8, 1, 0, 0, 12}, {16384, 16, 1, 0, 0, 12}, {16384, 2, 1, 0, 0, 12},
    {16384, 1, 0, 1, 0, 12}, {16384, 1, 0, 2, 0, 12}, {16384, 4, 1, 1, 0, 12}, {16384, 4, 1, 2, 0, 12},
    {16384, 1, 0, 0, 1, 12}, {16384, 4, 1, 0, 1, 12}, {14400, 1, 0, 0, 1, 12},
  };
  for (const V &v : vs) {
    const long long n_tiles = bytes / v.tile_bytes;
    for (int i = 0; i < 3; ++i) bulk_kernel<<<sms, v.warps * 32, v.warps * 16384>>>(buf, n_tiles, v.warps, v.tile_bytes, v.pieces, v.rot, v.map, v.depth);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) bulk_kernel<<<sms, v.warps * 32, v.warps * 16384>>>(buf, n_tiles, v.warps, v.tile_bytes, v.pieces, v.rot, v.map, v.depth);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("tile %5d B  pieces %2d rot %d  map %d  depth %d  warps %2d : %8.1f us  %8.1f GB/s\n", v.tile_bytes, v.pieces, v.rot, v.map, v.depth, v.warps,
           1e3f * ms / reps, (double)n_tiles * v.tile_bytes / (ms / reps * 1e-3) / 1e9);
  }
  unsigned *ctr;
  CK(cudaMalloc(&ctr, 4));
  CK(cudaFuncSetAttribute(ticket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 16384));
  for (int tb : {16384, 14400})
    for (unsigned chunk : {1u, 2u, 6u, 12u, 24u}) {
      const long long n_tiles = bytes / tb;
      float tot = 0;
      for (int i = 0; i < reps + 3; ++i) {
        CK(cudaMemsetAsync(ctr, 0, 4));
        CK(cudaEventRecord(e0));
        ticket_kernel<<<sms, 12 * 32, 12 * 16384>>>(buf, n_tiles, 12, tb, chunk, ctr);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 3) tot += ms;
      }
      printf("tickets: tile %5d B  chunk %2u windows : %8.1f us  %8.1f GB/s\n", tb, chunk, 1e3f * tot / reps, (double)n_tiles * tb / (tot / reps * 1e-3) / 1e9);
    }
  CK(cudaGetLastError());
  return 0;
}
