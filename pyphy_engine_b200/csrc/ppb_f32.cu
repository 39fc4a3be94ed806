// fp32 throughput kernels (FMA contraction allowed)
#include "ppb_launch_impl.cuh"
namespace ppb {
template struct Launch<float>;

cudaError_t launch_collision_flags(const float *traj, long long n_items, int n_steps, float thresh, uint8_t *flags,
                                   cudaStream_t st) {
  const long long total = n_items * (long long)(n_steps - 2);
  if (total <= 0) return cudaSuccess;
  collision_flag_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(traj, n_items, n_steps, thresh, flags);
  return cudaGetLastError();
}
}  // namespace ppb
