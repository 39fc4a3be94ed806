// ppb_post.cuh -- consumers of the frame / trajectory tensors, sm_100a.
//
//   crop_kernel        : the ball-centred, white-padded crops of DataSaver.fetch(cropSz)
//                        (dataio.py:160-182) for every (frame, world, ball), plus the re-based positions
//   collision_flag_kernel : utils.detect_collision's test |a|^2 >= thresh |v|^2 on the second
//                        differences of a recorded trajectory (utils.py:38-49)
#pragma once
#include "ppb_common.cuh"

namespace ppb {

__device__ __forceinline__ double round_half_away_d(double v) { return round(v); }   // Python-2 round()

// One CTA per (frame, world, ball).  frames: [n_frames][n_worlds][H][W][4] u8; traj: [n_frames][n_worlds][n_balls][2] (T);
// crops: [n_frames][n_worlds][n_balls][c][c][3] u8; pos_out: [n_frames][n_worlds][n_balls][2] f32.
// mode 0: DataSaver.fetch (dataio.py:160-179, offsets rounded half away from zero, positions re-based);
// mode 1: CustomWorld.get_glimpse (dataio.py:505-518, offsets truncated with int(), positions untouched).
template <typename T>
__global__ void crop_kernel(const uint8_t *frames, const T *traj, int n_worlds, int n_balls, int W, int H, int c, int mode,
                            uint8_t *crops, float *pos_out) {
  const long long item = blockIdx.x;   // (frame * n_worlds + world) * n_balls + ball
  const long long fw = item / n_balls;
  const double px = (double)traj[item * 2], py = (double)traj[item * 2 + 1];
  const double half = c / 2.0;
  const double xMid = round_half_away_d(px), yMid = round_half_away_d(py);
  double x1 = fmax(0.0, xMid - half), x2 = fmin((double)W, xMid + half);
  double y1 = fmax(0.0, yMid - half), y2 = fmin((double)H, yMid + half);
  int imX1, imX2, imY1, imY2, ix1, iy1;
  if (mode == 0) {
    imX1 = (int)round_half_away_d(half - (xMid - x1)); imX2 = (int)round_half_away_d(half + (x2 - xMid));
    imY1 = (int)round_half_away_d(half - (yMid - y1)); imY2 = (int)round_half_away_d(half + (y2 - yMid));
    ix1 = (int)round_half_away_d(x1); iy1 = (int)round_half_away_d(y1);
  } else {
    imX1 = (int)(half - (xMid - x1)); imX2 = (int)(half + (x2 - xMid));
    imY1 = (int)(half - (yMid - y1)); imY2 = (int)(half + (y2 - yMid));
    ix1 = (int)x1; iy1 = (int)y1;
  }
  const uint8_t *src = frames + (size_t)fw * H * W * 4;
  uint8_t *dst = crops + (size_t)item * c * c * 3;
  for (int i = threadIdx.x; i < c * c; i += blockDim.x) {
    const int cy = i / c, cx = i - cy * c;
    uint8_t r = 255, g = 255, b = 255;
    if (cy >= imY1 && cy < imY2 && cx >= imX1 && cx < imX2) {
      const int sy = iy1 + (cy - imY1), sx = ix1 + (cx - imX1);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const uchar4 p = reinterpret_cast<const uchar4 *>(src)[(size_t)sy * W + sx];
        r = p.x; g = p.y; b = p.z;
      }
    }
    dst[3 * i] = r; dst[3 * i + 1] = g; dst[3 * i + 2] = b;
  }
  if (threadIdx.x == 0 && pos_out) {
    const float fx = (float)px, fy = (float)py;
    if (mode == 0) {
      // position - x1 + imX1, evaluated on float32 values as the reference does on its float32 array
      pos_out[item * 2] = (float)(((double)fx - ix1) + imX1);
      pos_out[item * 2 + 1] = (float)(((double)fy - iy1) + imY1);
    } else {
      pos_out[item * 2] = fx;
      pos_out[item * 2 + 1] = fy;
    }
  }
}

// traj: [n_steps][n_items][2] float32 (what data.mat holds); flags: [n_items][n_steps - 2] u8,
// flags[i][k] = (|a_k|^2 >= thresh * |v_k|^2) with v = diff(pos), a = diff(v) in float32 (utils.py:42-48)
static __global__ void collision_flag_kernel(const float *traj, long long n_items, int n_steps, float thresh, uint8_t *flags) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = n_items * (long long)(n_steps - 2);
  if (i >= total) return;
  const long long it = i / (n_steps - 2);
  const int k = (int)(i - it * (n_steps - 2));
  const float2 p0 = reinterpret_cast<const float2 *>(traj)[(long long)k * n_items + it];
  const float2 p1 = reinterpret_cast<const float2 *>(traj)[(long long)(k + 1) * n_items + it];
  const float2 p2 = reinterpret_cast<const float2 *>(traj)[(long long)(k + 2) * n_items + it];
  const float v0x = __fsub_rn(p1.x, p0.x), v0y = __fsub_rn(p1.y, p0.y);
  const float v1x = __fsub_rn(p2.x, p1.x), v1y = __fsub_rn(p2.y, p1.y);
  const float ax = __fsub_rn(v1x, v0x), ay = __fsub_rn(v1y, v0y);
  const float a2 = __fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay));
  const float v2 = __fadd_rn(__fmul_rn(v0x, v0x), __fmul_rn(v0y, v0y));
  flags[i] = a2 >= __fmul_rn(thresh, v2) ? 1 : 0;
}

}  // namespace ppb
