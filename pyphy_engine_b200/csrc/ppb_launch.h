// ppb_launch.h -- host-callable launchers, one explicit instantiation per arithmetic
// (ppb_f32.cu is compiled with FMA contraction, ppb_f64.cu with -fmad=false).
#pragma once
#include "ppb_common.cuh"
#include "ppb_sampler.cuh"

namespace ppb {

template <typename T>
struct Launch {
  static size_t elem_size() { return sizeof(T); }
  static size_t hdr_size();
  // P.n_steps x Dynamics.step for every world in one launch (grid from advance_grid: persistent CTAs, worlds handed
  // out by tickets)
  static cudaError_t advance(const StatePtrs &S, const Layout &L, const StepParams &P, void *traj, int grid,
                             cudaStream_t st);
  static int advance_grid(int n_worlds, int n_balls, int sm_count);
  // WallRC of worlds [world0, world0 + count)
  static cudaError_t wall_cache(const StatePtrs &S, const Layout &L, const RenderLayout &RL, int world0, int count,
                                cudaStream_t st);
  static cudaError_t render(const StatePtrs &S, const Layout &L, const RenderLayout &RL, const uint32_t *atlas,
                            uint8_t *frames, cudaStream_t st);
  static cudaError_t reset_cache(const StatePtrs &S, int n_balls, int world0, int count, int keep_cache,
                                 int32_t step_value, cudaStream_t st);
  static cudaError_t sample_rect(const StatePtrs &S, int n_balls, int n_walls, int world0, int count,
                                 const SamplerParams &P, int32_t step_value, int32_t *fail_count, cudaStream_t st);
  static cudaError_t crop(const uint8_t *frames, const void *traj, long long n_frames, int n_worlds, int n_balls, int W,
                          int H, int c, int mode, uint8_t *crops, float *pos_out, cudaStream_t st);
  // fetch(cropSz, procSz): PIL-bilinear resize of crops [items][c][c][3] -> [items][p][p][3], positions scaled in place,
  // displacements between consecutive frames scaled into vel_out (may be NULL)
  static cudaError_t resize_crops(const uint8_t *crops, const void *traj, long long n_frames, long long items_per_frame, int c,
                                  int p, uint8_t *out, float *pos, float *vel_out, cudaStream_t st);
  // DynamicsHorizon look-ahead matrices of every row of a trajectory [n_rows][n_items][2]
  static cudaError_t horizon(const void *traj, long long n_items, int n_rows, int lookahead, float *out, cudaStream_t st);
  // Dynamics.apply_force (primitives.py:580-594): force is a device array [count][n_balls][2] of fp64
  static cudaError_t apply_force(const StatePtrs &S, int n_balls, int world0, int count, const double *force,
                                 double force_t, cudaStream_t st);
  // out[0..2] = {min floor(radius), max ceil(radius), any non-integer radius} over the occupied slots (out pre-set)
  static cudaError_t radius_range(const StatePtrs &S, int n_walls, int n_balls, int32_t *out, cudaStream_t st);
  // traj (batch dtype) -> float32
  static cudaError_t traj_to_f32(const void *src, float *dst, long long n, cudaStream_t st);
};

cudaError_t launch_collision_flags(const float *traj, long long n_items, int n_steps, float thresh, uint8_t *flags,
                                   cudaStream_t st);
cudaError_t launch_build_sprites(uint32_t *atlas, const RenderLayout &RL, int n_balls, const uint32_t *fill8,
                                 const uint32_t *stroke8, cudaStream_t st);

}  // namespace ppb
