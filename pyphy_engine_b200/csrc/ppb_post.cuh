// ppb_post.cuh -- consumers of the frame / trajectory tensors, sm_100a.
//
//   crop_kernel        : the ball-centred, white-padded crops of DataSaver.fetch(cropSz)
//                        (dataio.py:160-182) for every (frame, world, ball), plus the re-based positions
//   resize_crops_kernel : the procSz step of fetch (dataio.py:183-189): scipy.misc.imresize(imBall, (procSz, procSz))
//                        is PIL's Image.resize(..., BILINEAR); this is Pillow's 8-bit two-pass resampler
//                        (libImaging/Resample.c: triangle filter widened by the scale factor, 22-bit fixed-point
//                        coefficients, horizontal pass rounded to 8 bits, then the vertical pass), plus the
//                        position / displacement rescaling of dataio.py:184-189
//   horizon_kernel     : DynamicsHorizon.get_data's look-ahead matrix (primitives.py:839-852) for every frame of a
//                        recorded trajectory at once
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

// ---- Pillow's precompute_coeffs + normalize_coeffs_8bpc for output index xx (in0 = 0, in1 = in_size, bilinear) ----
// bounds: xmin, count; k[0..count): coefficients scaled by 2^22.  Explicitly rounded double operations: no FMA
// contraction may move a coefficient by one unit.
#define PPB_RESAMPLE_BITS 22
__device__ inline void bilinear_coeffs(int in_size, int out_size, int xx, int ksize, int *xmin_out, int *count_out, int *k) {
  double scale = __ddiv_rn((double)in_size, (double)out_size), filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = filterscale;                        // BILINEAR.support = 1.0
  const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
  const double ss = __ddiv_rn(1.0, filterscale);
  int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
  if (xmin < 0) xmin = 0;
  int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double w[32];
  double ww = 0.0;
  for (int x = 0; x < xmax && x < 32; ++x) {
    double t = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
    if (t < 0.0) t = -t;
    w[x] = t < 1.0 ? __dsub_rn(1.0, t) : 0.0;
    ww = __dadd_rn(ww, w[x]);
  }
  for (int x = 0; x < ksize; ++x) {
    double v = 0.0;
    if (x < xmax && x < 32) v = ww != 0.0 ? __ddiv_rn(w[x], ww) : w[x];
    const double sc = __dmul_rn(v, (double)(1 << PPB_RESAMPLE_BITS));
    k[x] = v < 0.0 ? (int)__dadd_rn(-0.5, sc) : (int)__dadd_rn(0.5, sc);
  }
  *xmin_out = xmin;
  *count_out = xmax;
}
__device__ __forceinline__ uint8_t resample_clip8(int v) {
  v >>= PPB_RESAMPLE_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One CTA per (frame, world, ball): crops [items][c][c][3] -> out [items][p][p][3]; positions (crop coordinates,
// float32, in place) times p/c; vel_out[frame i-1] = (centre_i - centre_{i-1}) * p/c from the recorded trajectory.
// Dynamic shared memory: bounds 2p ints, coefficients p*ksize ints, the crop 3c^2 bytes, the horizontal pass 3cp bytes.
template <typename T>
__global__ void resize_crops_kernel(const uint8_t *crops, const T *traj, long long items_per_frame, int c, int p, int ksize,
                                    uint8_t *out, float *pos, float *vel_out) {
  extern __shared__ __align__(16) unsigned char rs_smem[];
  int *bounds = reinterpret_cast<int *>(rs_smem);
  int *kk = bounds + 2 * p;
  uint8_t *src = reinterpret_cast<uint8_t *>(kk + p * ksize);
  uint8_t *tmp = src + 3 * c * c;
  const long long item = blockIdx.x;
  for (int xx = threadIdx.x; xx < p; xx += blockDim.x) bilinear_coeffs(c, p, xx, ksize, &bounds[2 * xx], &bounds[2 * xx + 1], kk + xx * ksize);
  const uint8_t *g = crops + (size_t)item * c * c * 3;
  for (int i = threadIdx.x; i < 3 * c * c; i += blockDim.x) src[i] = g[i];
  __syncthreads();
  // the rows the vertical pass reads (Pillow resamples only those horizontally; the others are never used)
  for (int i = threadIdx.x; i < c * p; i += blockDim.x) {
    const int y = i / p, xx = i - y * p;
    const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
    const int *k = kk + xx * ksize;
    int s0 = 1 << (PPB_RESAMPLE_BITS - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
      const uint8_t *q = src + (y * c + x + xmin) * 3;
      s0 += q[0] * k[x]; s1 += q[1] * k[x]; s2 += q[2] * k[x];
    }
    uint8_t *d = tmp + (y * p + xx) * 3;
    d[0] = resample_clip8(s0); d[1] = resample_clip8(s1); d[2] = resample_clip8(s2);
  }
  __syncthreads();
  uint8_t *o = out + (size_t)item * p * p * 3;
  for (int i = threadIdx.x; i < p * p; i += blockDim.x) {
    const int yy = i / p, xx = i - yy * p;
    const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
    const int *k = kk + yy * ksize;
    int s0 = 1 << (PPB_RESAMPLE_BITS - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < n; ++y) {
      const uint8_t *q = tmp + ((y + ymin) * p + xx) * 3;
      s0 += q[0] * k[y]; s1 += q[1] * k[y]; s2 += q[2] * k[y];
    }
    o[3 * i] = resample_clip8(s0); o[3 * i + 1] = resample_clip8(s1); o[3 * i + 2] = resample_clip8(s2);
  }
  if (threadIdx.x < 2) {
    const double pos_scale = __ddiv_rn((double)p, (double)c);          // float(procSz) / float(cropSz), dataio.py:184
    const int ch = threadIdx.x;
    if (pos) pos[item * 2 + ch] = (float)__dmul_rn((double)pos[item * 2 + ch], pos_scale);   // float32 array element * float
    if (vel_out) {
      const long long frame = item / items_per_frame;
      if (frame > 0) {                                                 // vx = pos.x() - pPos[j][0]; velocity[.., i-1] = vx * posScale
        const double dv = __dsub_rn((double)traj[item * 2 + ch], (double)traj[(item - items_per_frame) * 2 + ch]);
        vel_out[(item - items_per_frame) * 2 + ch] = (float)__dmul_rn(dv, pos_scale);
      }
    }
  }
}

// DynamicsHorizon.get_data (primitives.py:839-852), all frames at once.  traj: [n_rows][n_items][2] (row 0 = the
// positions before the first step); out: [n_rows - N][n_items][2 N] float32, out[t][i][2 k + c] = the displacement of
// item i during step t + k + 1, i.e. float32(traj[t + k + 1] - traj[t + k]).
template <typename T>
__global__ void horizon_kernel(const T *traj, long long n_items, int n_rows, int N, float *out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(n_rows - N) * n_items * N;
  if (i >= total) return;
  const int k = (int)(i % N);
  const long long it = (i / N) % n_items;
  const long long t = i / ((long long)N * n_items);
  const T *a = traj + ((t + k) * n_items + it) * 2, *b = a + n_items * 2;
  float2 d;
  d.x = (float)(b[0] - a[0]);
  d.y = (float)(b[1] - a[1]);
  reinterpret_cast<float2 *>(out)[i] = d;
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
