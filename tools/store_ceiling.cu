// store_ceiling.cu -- what can a store-only kernel reach on this GPU?  The render kernel writes 16 KiB frames
// and reads almost nothing, so the copy bandwidth in MEASURED_PEAKS.json (read + write) is not its natural
// ceiling.  Three store-only variants over the same 2 GiB the bench's render launch writes:
//   memset      cudaMemsetAsync
//   st.v4       grid-stride 16-byte stores from registers
//   bulk 16K    the render kernel's skeleton without the drawing: one persistent CTA per SM, every warp
//               owns a 16 KiB shared-memory tile and issues cp.async.bulk.global.shared::cta per frame
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/store_ceiling tools/store_ceiling.cu
//   tools/_bin/store_ceiling [bytes]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void st_v4_kernel(uint4 *dst, long long n16) {
  const uint4 v = make_uint4(0xff0000ffu, 0xff0000ffu, 0xffffffffu, 0xffffffffu);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = v;
}

template <int TILE_BYTES>
__global__ void bulk_kernel(uint8_t *dst, long long n_tiles, int warps) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *tile = smem + (size_t)warp * TILE_BYTES;
  for (int i = lane * 16; i < TILE_BYTES; i += 32 * 16) *reinterpret_cast<uint4 *>(tile + i) = make_uint4(0xff0000ffu, ~0u, ~0u, ~0u);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
  const long long stride = (long long)gridDim.x * warps;
  for (long long t = (long long)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * TILE_BYTES), "r"(saddr), "n"(TILE_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the tile may be rewritten (as the renderer does)
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the same skeleton with a 2-D tensor-map store (frames as an array of 128-byte lines, box = 32 pixels x 128 lines):
// what the copy engine's shared-memory swizzle modes cost
__global__ void tensor_kernel(const __grid_constant__ CUtensorMap tmap, long long n_tiles, int warps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *tile = smem + (size_t)warp * 16384;
  for (int i = lane * 16; i < 16384; i += 32 * 16) *reinterpret_cast<uint4 *>(tile + i) = make_uint4(0xff0000ffu, ~0u, ~0u, ~0u);
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
  const long long stride = (long long)gridDim.x * warps;
  for (long long t = (long long)blockIdx.x * warps + warp; t < n_tiles; t += stride) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tmap), "r"(0), "r"((int)(t * 128)), "r"(saddr) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
  const long long bytes = argc > 1 ? atoll(argv[1]) : (2LL << 30);
  uint8_t *buf;
  CK(cudaMalloc(&buf, bytes));
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int reps = 20;
  auto report = [&](const char *name, float ms) { printf("%-28s %8.1f us  %8.1f GB/s\n", name, 1e3f * ms / reps, bytes / (ms / reps * 1e-3) / 1e9); };
  float ms;
  // memset
  for (int i = 0; i < 3; ++i) CK(cudaMemsetAsync(buf, 0xff, bytes));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) CK(cudaMemsetAsync(buf, 0xff, bytes));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  report("cudaMemsetAsync", ms);
  // st.v4
  for (int cta_per_sm : {2, 4, 8}) {
    for (int i = 0; i < 3; ++i) st_v4_kernel<<<sms * cta_per_sm, 256>>>((uint4 *)buf, bytes / 16);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) st_v4_kernel<<<sms * cta_per_sm, 256>>>((uint4 *)buf, bytes / 16);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[64];
    snprintf(nm, sizeof nm, "st.global.v4 %d CTA/SM", cta_per_sm);
    report(nm, ms);
  }
  // bulk stores of 16 KiB tiles
  CK(cudaFuncSetAttribute(bulk_kernel<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 13 * 16384));
  for (int warps : {4, 8, 12, 13}) {
    const long long n_tiles = bytes / 16384;
    for (int i = 0; i < 3; ++i) bulk_kernel<16384><<<sms, warps * 32, warps * 16384>>>(buf, n_tiles, warps);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) bulk_kernel<16384><<<sms, warps * 32, warps * 16384>>>(buf, n_tiles, warps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[64];
    snprintf(nm, sizeof nm, "bulk 16 KiB x %d warps/SM", warps);
    report(nm, ms);
  }
  // two tiles per bulk copy
  CK(cudaFuncSetAttribute(bulk_kernel<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768));
  for (int warps : {4, 6}) {
    const long long n_tiles = bytes / 32768;
    for (int i = 0; i < 3; ++i) bulk_kernel<32768><<<sms, warps * 32, warps * 32768>>>(buf, n_tiles, warps);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) bulk_kernel<32768><<<sms, warps * 32, warps * 32768>>>(buf, n_tiles, warps);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    char nm[64];
    snprintf(nm, sizeof nm, "bulk 32 KiB x %d warps/SM", warps);
    report(nm, ms);
  }
  // tensor-map stores, by swizzle mode
  {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    encode_fn encode = (encode_fn)fn;
    CK(cudaFuncSetAttribute(tensor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384));
    const CUtensorMapSwizzle modes[2] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_128B};   // (32B / 64B: the 128-byte box row exceeds the span)
    const char *names[2] = {"none", "128B"};
    for (int m = 0; m < 2; ++m) {
      CUtensorMap tmap;
      const cuuint64_t dims[2] = {32, (cuuint64_t)(bytes / 128)};
      const cuuint64_t strides[1] = {128};
      const cuuint32_t box[2] = {32, 128};
      const cuuint32_t estr[2] = {1, 1};
      if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      const int warps = 12;
      const long long n_tiles = bytes / 16384;
      for (int i = 0; i < 3; ++i) tensor_kernel<<<sms, warps * 32, warps * 16384>>>(tmap, n_tiles, warps);
      CK(cudaEventRecord(e0));
      for (int i = 0; i < reps; ++i) tensor_kernel<<<sms, warps * 32, warps * 16384>>>(tmap, n_tiles, warps);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      char nm[64];
      snprintf(nm, sizeof nm, "tensor 16 KiB swizzle %s", names[m]);
      report(nm, ms);
    }
  }
  CK(cudaGetLastError());
  CK(cudaFree(buf));
  return 0;
}
