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
cudaError_t launch_build_sprites(uint32_t *atlas, const RenderLayout &RL, int n_balls, const uint32_t *fill8,
                                 const uint32_t *stroke8, cudaStream_t st) {
  const int nr = RL.r_max - RL.r_min + 1;
  int max_sz = 0;
  for (int b = 0; b < n_balls; ++b) {
    const int sz = 2 * (RL.r_max + RL.s_thick[b]);
    if (sz > max_sz) max_sz = sz;
  }
  if (nr <= 0 || max_sz <= 0) return cudaSuccess;
  dim3 grid((max_sz * max_sz + 127) / 128, n_balls, nr);
  build_ball_sprites_kernel<<<grid, 128, 0, st>>>(atlas, RL, n_balls, fill8, stroke8);
  return cudaGetLastError();
}
}  // namespace ppb
