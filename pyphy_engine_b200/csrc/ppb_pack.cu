// ppb_pack.cu -- frames leave the device run-length packed (ppb_simulate_host*: World.generate_image results
// delivered into the caller's host array, the way DataSaver.save_sequence consumes them, dataio.py:118-128).
//
// A rendered canvas is opaque (alpha is 255 everywhere: OVER onto an opaque white canvas, primitives.py:876, 998)
// and mostly flat colour, so a frame is sent as 32-bit tokens  rgb | run_length << 24  (runs end at every 64th
// pixel): about 2.4 KB instead of 16 KB for a 64x64 billiards frame.  The download then costs PCIe a seventh of the
// raw copy and the RGBA pixels are rebuilt in the caller's array by a pool of host threads with streaming stores
// (one pass over host memory, like the DMA engine's own write).  Frames that are not opaque, or that do not
// compress to a quarter of their size, travel raw.
//
//   device : pack_frames_kernel, one warp per frame, two passes (count, then emit at an atomically claimed offset)
//   host   : PackPipe -- two job slots; a coordinator thread waits for the packing, starts the token copy and
//            hands the unpacking to the worker pool; ppb_host_sync waits on the job counter
#include <emmintrin.h>
#include <sched.h>
#include <unistd.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ppb_pack.h"

namespace ppb {
namespace {

// ctrl[0] = tokens claimed so far, ctrl[1] = flags (1: a pixel is not opaque, 2: the token buffer is too small)
__global__ void __launch_bounds__(256) pack_frames_kernel(const uint32_t *frames, long long n_frames, int npx,
                                                          uint32_t *tokens, uint32_t cap, uint2 *table, uint32_t *ctrl) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long fr = warp0; fr < n_frames; fr += n_warps) {
    const uint32_t *f = frames + fr * npx;
    // a 64-pixel group: lane holds pixels base + lane and base + 32 + lane; bit p of M = "a run starts at pixel p"
    auto group = [&](int base, uint32_t &a, uint32_t &b, bool &sa, bool &sb) -> unsigned long long {
      const int p0 = base + lane, p1 = p0 + 32;
      a = p0 < npx ? f[p0] : 0u;
      b = p1 < npx ? f[p1] : 0u;
      const uint32_t prev_a = __shfl_up_sync(0xffffffffu, a, 1), last_a = __shfl_sync(0xffffffffu, a, 31);
      const uint32_t prev_b = __shfl_up_sync(0xffffffffu, b, 1);
      sa = p0 < npx && (lane == 0 || a != prev_a);
      sb = p1 < npx && (lane == 0 ? b != last_a : b != prev_b);
      return (unsigned long long)__ballot_sync(0xffffffffu, sa) | ((unsigned long long)__ballot_sync(0xffffffffu, sb) << 32);
    };
    uint32_t n = 0;
    bool bad = false;
    for (int base = 0; base < npx; base += 64) {
      uint32_t a, b;
      bool sa, sb;
      n += (uint32_t)__popcll(group(base, a, b, sa, sb));
      bad |= (base + lane < npx && (a >> 24) != 255u) || (base + 32 + lane < npx && (b >> 24) != 255u);
    }
    bad = __any_sync(0xffffffffu, bad);
    uint32_t off = 0;
    if (lane == 0) {
      off = atomicAdd(ctrl, n);
      uint32_t flags = bad ? 1u : 0u;
      if ((unsigned long long)off + n > cap) flags |= 2u;
      if (flags) atomicOr(ctrl + 1, flags);
      table[fr] = make_uint2(off, n);
    }
    off = __shfl_sync(0xffffffffu, off, 0);
    if (bad || (unsigned long long)off + n > cap) continue;
    uint32_t w = off;
    for (int base = 0; base < npx; base += 64) {
      uint32_t a, b;
      bool sa, sb;
      const unsigned long long M = group(base, a, b, sa, sb);
      const int valid = min(64, npx - base);
      if (sa) {
        const int p = lane;
        const unsigned long long rest = M >> (p + 1);
        const int cnt = rest ? __ffsll((long long)rest) : valid - p;
        tokens[w + __popcll(M & ((1ull << p) - 1ull))] = (a & 0x00ffffffu) | ((uint32_t)cnt << 24);
      }
      if (sb) {
        const int p = 32 + lane;
        const unsigned long long rest = p + 1 < 64 ? (M >> (p + 1)) : 0ull;
        const int cnt = rest ? __ffsll((long long)rest) : valid - p;
        tokens[w + __popcll(M & ((1ull << p) - 1ull))] = (b & 0x00ffffffu) | ((uint32_t)cnt << 24);
      }
      w += (uint32_t)__popcll(M);
    }
  }
}

// ---- host side: tokens -> RGBA --------------------------------------------------------------------------------------
constexpr int kBlockPx = 4096;   // pixels decoded into the cache-resident block before they are streamed out

inline void stream_out(uint8_t *dst, const uint32_t *src, size_t n_px) {
  const size_t bytes = n_px * 4;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && (bytes & 63u) == 0) {
    const __m128i *s = reinterpret_cast<const __m128i *>(src);
    __m128i *d = reinterpret_cast<__m128i *>(dst);
    for (size_t i = 0; i < bytes / 16; i += 4) {
      _mm_stream_si128(d + i, _mm_load_si128(s + i));
      _mm_stream_si128(d + i + 1, _mm_load_si128(s + i + 1));
      _mm_stream_si128(d + i + 2, _mm_load_si128(s + i + 2));
      _mm_stream_si128(d + i + 3, _mm_load_si128(s + i + 3));
    }
  } else {
    memcpy(dst, src, bytes);
  }
}

// The 64x64 canvas (any frame of at most kBlockPx pixels whose byte size is a multiple of 64) with 256-bit stores:
// one unconditional store covers a run of up to 8 pixels, the block leaves with 32-byte streaming stores.
__attribute__((target("avx2"))) bool unpack_small_avx2(const uint32_t *tokens, const uint2 *table, long long f0, long long f1,
                                                       int npx, uint8_t *frames_host) {
  alignas(64) uint32_t block[kBlockPx + 128];
  for (long long fr = f0; fr < f1; ++fr) {
    const uint32_t *t = tokens + table[fr].x, *te = t + table[fr].y;
    uint32_t *p = block;
    while (t < te) {
      const uint32_t tok = *t++;
      const int cnt = (int)(tok >> 24);
      const __m256i v = _mm256_set1_epi32((int)(tok | 0xff000000u));
      _mm256_storeu_si256(reinterpret_cast<__m256i *>(p), v);
      if (cnt > 8)
        for (int i = 8; i < cnt; i += 8) _mm256_storeu_si256(reinterpret_cast<__m256i *>(p + i), v);
      p += cnt;
      if (p > block + kBlockPx) return false;
    }
    if (p - block != npx) return false;
    const __m256i *s = reinterpret_cast<const __m256i *>(block);
    __m256i *d = reinterpret_cast<__m256i *>(frames_host + (size_t)fr * npx * 4);
    for (int i = 0; i < npx / 8; i += 2) {
      _mm256_stream_si256(d + i, _mm256_load_si256(s + i));
      _mm256_stream_si256(d + i + 1, _mm256_load_si256(s + i + 1));
    }
  }
  _mm_sfence();
  return true;
}

// The same with 512-bit registers: two unconditional 64-byte stores cover a run of up to 32 pixels (the branch for the
// longer runs -- white rows without a ball -- is rare and predictable), and the block leaves as whole cache lines.
// Two blocks: while frame k is decoded into one, frame k-1 is streamed out of the other, one line per two tokens, so
// the streaming stores (bound by the write-combining buffers) and the decoder's L1 stores overlap.
__attribute__((target("avx512f"))) bool unpack_small_avx512(const uint32_t *tokens, const uint2 *table, long long f0,
                                                            long long f1, int npx, uint8_t *frames_host) {
  alignas(64) uint32_t blocks[2][kBlockPx + 128];
  const int n_lines = npx / 16;
  uint8_t *prev_dst = nullptr;   // where the frame decoded last goes
  int prev_line = 0;             // lines of it already streamed
  int cur = 0;
#define PPB_STREAM_LINE(src)                                                                          \
  do {                                                                                                \
    _mm512_stream_si512(reinterpret_cast<__m512i *>(prev_dst + (size_t)prev_line * 64),               \
                        _mm512_load_si512((src) + prev_line * 16));                                   \
    ++prev_line;                                                                                      \
  } while (0)
  bool ok = true;
  for (long long fr = f0; fr < f1 && ok; ++fr) {
    uint32_t *block = blocks[cur];
    const uint32_t *pblock = blocks[cur ^ 1];
    const uint32_t *t = tokens + table[fr].x, *te = t + table[fr].y;
    uint32_t *p = block;
    if ((long long)table[fr].y > npx) { ok = false; break; }   // every token holds at least one pixel
    unsigned k = 0;
    while (t < te) {
      const uint32_t tok = *t++;
      const int cnt = (int)(tok >> 24);
      const __m512i v = _mm512_set1_epi32((int)(tok | 0xff000000u));
      _mm512_storeu_si512(p, v);
      _mm512_storeu_si512(p + 16, v);
      if (cnt > 32) {
        if (cnt > 64) { ok = false; break; }
        _mm512_storeu_si512(p + 32, v);
        _mm512_storeu_si512(p + 48, v);
      }
      p += cnt;
      if (p > block + kBlockPx) { ok = false; break; }
      if ((++k & 1u) && prev_dst && prev_line < n_lines) PPB_STREAM_LINE(pblock);
    }
    if (p - block != npx) ok = false;
    if (prev_dst)
      while (prev_line < n_lines) PPB_STREAM_LINE(pblock);
    prev_dst = ok ? frames_host + (size_t)fr * npx * 4 : nullptr;
    prev_line = 0;
    cur ^= 1;
  }
  if (prev_dst)
    while (prev_line < n_lines) PPB_STREAM_LINE(blocks[cur ^ 1]);
#undef PPB_STREAM_LINE
  _mm_sfence();
  return ok;
}

// frames [f0, f1) of one job; returns false when a frame's tokens do not add up to its pixel count
bool unpack_range(const uint32_t *tokens, const uint2 *table, long long f0, long long f1, int npx, uint8_t *frames_host) {
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  static const bool have_avx512 = __builtin_cpu_supports("avx512f") && !getenv("PPB_NO_AVX512");
  if (have_avx512 && npx <= kBlockPx && (npx & 31) == 0 && (reinterpret_cast<uintptr_t>(frames_host) & 63u) == 0)
    return unpack_small_avx512(tokens, table, f0, f1, npx, frames_host);
  if (have_avx2 && npx <= kBlockPx && (npx & 15) == 0 && (reinterpret_cast<uintptr_t>(frames_host) & 31u) == 0)
    return unpack_small_avx2(tokens, table, f0, f1, npx, frames_host);
  alignas(64) uint32_t block[kBlockPx + 128];
  for (long long fr = f0; fr < f1; ++fr) {
    const uint32_t *t = tokens + table[fr].x, *te = t + table[fr].y;
    uint8_t *dst = frames_host + (size_t)fr * npx * 4;
    long long done = 0;   // pixels of this frame already streamed out
    int fill = 0;         // pixels in the block
    while (t < te) {
      const uint32_t tok = *t++;
      const int cnt = (int)(tok >> 24);
      const __m128i v = _mm_set1_epi32((int)(tok | 0xff000000u));
      uint32_t *p = block + fill;
      // two unconditional stores cover the short runs (most tokens) without a data-dependent branch; what they write
      // beyond the run is overwritten by the runs that follow
      _mm_storeu_si128(reinterpret_cast<__m128i *>(p), v);
      _mm_storeu_si128(reinterpret_cast<__m128i *>(p + 4), v);
      if (cnt > 8)
        for (int i = 8; i < cnt; i += 4) _mm_storeu_si128(reinterpret_cast<__m128i *>(p + i), v);
      fill += cnt;
      if (fill >= kBlockPx) {                       // a run is at most 64 pixels: the block has room for the excess
        stream_out(dst + (size_t)done * 4, block, kBlockPx);
        done += kBlockPx;
        fill -= kBlockPx;
        memcpy(block, block + kBlockPx, (size_t)fill * 4);
      }
    }
    if (done + fill != npx) return false;
    if (fill) memcpy(dst + (size_t)done * 4, block, (size_t)fill * 4);
  }
  _mm_sfence();
  return true;
}

#ifdef PPB_PACK_HOST_CHECK
}  // namespace
// tools/unpack_host_check.cpp: the host half alone (no GPU needed)
bool host_check_unpack(const uint32_t *tokens, const uint2 *table, long long f0, long long f1, int npx, uint8_t *out) {
  return unpack_range(tokens, table, f0, f1, npx, out);
}
namespace {
#endif

constexpr int kMaxPieces = 4;   // a job's packed frames are packed, copied and rebuilt in up to this many pieces

// One ppb_simulate_host* chunk / ppb_download_frames piece.  Frames [0, n_raw) travel raw through the copy engine;
// frames [n_raw, n_frames) are packed in `n_pieces` pieces, so that the host threads start on the first piece while
// the tokens of the next ones still cross PCIe (a step's download starts ~1.5 ms instead of ~6 ms after its kernels).
struct Job {
  int slot = 0;
  const uint8_t *frames_dev = nullptr;
  uint8_t *frames_host = nullptr;
  long long n_frames = 0;
  long long n_raw = 0;
  int npx = 0;
  int n_pieces = 0;
  long long piece_f0[kMaxPieces + 1] = {0};   // piece c = packed frames [piece_f0[c], piece_f0[c + 1]) (relative to n_raw)
  size_t piece_tok0[kMaxPieces + 1] = {0};    // its region of the slot's token buffer
  // progress (PackPipe::mu)
  int pending = 0;            // pieces not yet rebuilt + 1 for the coordinator
  bool ok = true;
  double s_tok = 0, s_raw = 0, s_unp = 0;   // seconds: token copies, raw copy, rebuilding (slowest thread, summed over pieces)
  long long frames_unpacked = 0;
};

}  // namespace

struct PackPipe {
  int device = 0;
  unsigned n_workers = 1;
  // per slot
  uint32_t *d_tokens[2] = {nullptr, nullptr};
  uint2 *d_table[2] = {nullptr, nullptr};
  uint32_t *d_ctrl[2] = {nullptr, nullptr};    // 4 words per piece
  uint32_t *h_tokens[2] = {nullptr, nullptr};
  uint2 *h_table[2] = {nullptr, nullptr};
  uint32_t *h_ctrl[2] = {nullptr, nullptr};
  size_t cap_tokens[2] = {0, 0}, cap_frames[2] = {0, 0};
  cudaEvent_t ev_packed[2][kMaxPieces] = {};
  cudaStream_t copy_stream = nullptr;
  // job flow: submit() -> coordinator (copies) -> host threads (rebuild) -> done
  std::mutex mu;
  std::condition_variable cv_coord, cv_work, cv_done;
  std::deque<Job *> q_coord;
  struct Work {   // one piece of a job, rebuilt by all host threads together
    Job *job = nullptr;
    int piece = 0;
    long long id = 0;
    std::atomic<long long> next{0};
    int left = 0;
    bool ok = true, started = false;
    std::chrono::steady_clock::time_point t_start;
  };
  std::deque<Work *> q_work;
  long long submitted = 0, finished = 0, next_work_id = 0;
  bool slot_busy[2] = {false, false};
  bool stop = false;
  std::string error;
  long long jobs_raw = 0, jobs_packed = 0;
  unsigned long long bytes_packed = 0, bytes_raw = 0;
  double us_copy = 0, us_unpack = 0;   // time in the token copies / in the rebuilding, summed over the jobs
  // Split of a job between the copy engine (raw frames) and the host threads (packed frames): both write the caller's
  // array at the same time, so a job takes max(token copies + raw copy, rebuilding) and the split that equalises the
  // two is tracked from the measured costs (seconds per frame; 0: not measured yet).  On the boxes measured the copy
  // engine reaches ~27 GB/s next to 16 streaming threads (52-57 GB/s alone), the threads 115-140 GB/s.
  double raw_fraction = 0.1, c_raw = 0, c_tok = 0, c_unp = 0;
  bool split_fixed = false;
  std::thread coordinator;
  std::vector<std::thread> workers;

  // mu held.  One party of a job (a rebuilt piece, or the coordinator when its copies are done) is finished.
  void job_part_done(Job *j, bool ok, const char *err) {
    if (!ok) {
      j->ok = false;
      if (error.empty()) error = err;
    }
    if (--j->pending > 0) return;
    us_copy += 1e6 * j->s_tok;
    us_unpack += 1e6 * j->s_unp;
    update_split(j);
    slot_busy[j->slot] = false;
    ++finished;
    delete j;
    cv_done.notify_all();
  }

  // mu held
  void update_split(const Job *j) {
    auto ema = [](double &c, double v) { c = c > 0 ? 0.5 * c + 0.5 * v : v; };
    if (j->frames_unpacked > 0) {
      ema(c_tok, j->s_tok / (double)j->frames_unpacked);
      ema(c_unp, j->s_unp / (double)j->frames_unpacked);
    }
    if (j->n_raw > 0 && j->s_raw > 0) ema(c_raw, j->s_raw / (double)j->n_raw);
    static const bool debug = getenv("PPB_PACK_DEBUG") != nullptr;
    if (debug)
      fprintf(stderr, "[pack] raw %lld frames %.2f ms (%.1f GB/s), tokens %.2f ms, rebuilt %lld frames %.2f ms (%.1f GB/s); split %.3f\n",
              j->n_raw, 1e3 * j->s_raw, j->s_raw > 0 ? j->n_raw * j->npx * 4e-9 / j->s_raw : 0.0, 1e3 * j->s_tok,
              j->frames_unpacked, 1e3 * j->s_unp, j->s_unp > 0 ? j->frames_unpacked * j->npx * 4e-9 / j->s_unp : 0.0, raw_fraction);
    if (split_fixed || c_unp <= 0 || c_raw <= 0) return;
    // n f c_raw + n (1 - f) c_tok = n (1 - f) c_unp
    const double gain = c_unp - c_tok;
    double f = gain > 0 ? gain / (c_raw + gain) : 0.0;
    f = 0.75 * raw_fraction + 0.25 * f;
    raw_fraction = f < 0.0 ? 0.0 : (f > 0.6 ? 0.6 : f);
  }

  void coordinator_loop() {
    cudaSetDevice(device);
    for (;;) {
      Job *j = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_coord.wait(lk, [&] { return stop || !q_coord.empty(); });
        if (q_coord.empty()) return;
        j = q_coord.front();
        q_coord.pop_front();
      }
      const int s = j->slot;
      const size_t frame_bytes = (size_t)j->npx * 4;
      const uint8_t *packed_dev = j->frames_dev + (size_t)j->n_raw * frame_bytes;
      uint8_t *packed_host = j->frames_host + (size_t)j->n_raw * frame_bytes;
      bool ok = true;
      const char *err = nullptr;
      bool any_packed = false;
      int handled = 0;   // pieces handed to the host threads or finished here
      for (int c = 0; c < j->n_pieces && ok; ++c) {
        if (cudaEventSynchronize(ev_packed[s][c]) != cudaSuccess) { ok = false; err = "pack kernel failed"; break; }
        const uint32_t total = h_ctrl[s][4 * c], flags = h_ctrl[s][4 * c + 1];
        const long long f0 = j->piece_f0[c], nf = j->piece_f0[c + 1] - f0;
        const auto t0 = std::chrono::steady_clock::now();
        if (flags) {   // not opaque, or not compressible: the piece travels as it is
          cudaError_t e = cudaMemcpyAsync(packed_host + (size_t)f0 * frame_bytes, packed_dev + (size_t)f0 * frame_bytes,
                                          (size_t)nf * frame_bytes, cudaMemcpyDeviceToHost, copy_stream);
          if (e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
          if (e != cudaSuccess) { ok = false; err = "raw frame copy failed"; }
          std::lock_guard<std::mutex> lk(mu);
          bytes_raw += (size_t)nf * frame_bytes;
          ++handled;
          job_part_done(j, true, nullptr);   // nothing to rebuild for this piece (pending stays >= 1: the coordinator's own part)
          continue;
        }
        cudaError_t e = cudaMemcpyAsync(h_tokens[s] + j->piece_tok0[c], d_tokens[s] + j->piece_tok0[c], (size_t)total * 4,
                                        cudaMemcpyDeviceToHost, copy_stream);
        if (e == cudaSuccess)
          e = cudaMemcpyAsync(h_table[s] + f0, d_table[s] + f0, (size_t)nf * sizeof(uint2), cudaMemcpyDeviceToHost, copy_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
        if (e != cudaSuccess) { ok = false; err = "token copy failed"; break; }
        any_packed = true;
        Work *w = new Work;
        w->job = j;
        w->piece = c;
        w->left = (int)n_workers;
        {
          std::lock_guard<std::mutex> lk(mu);
          w->id = next_work_id++;
          j->s_tok += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
          bytes_packed += (size_t)total * 4 + (size_t)nf * sizeof(uint2);
          q_work.push_back(w);
          ++handled;
        }
        cv_work.notify_all();
      }
      if (handled < j->n_pieces) {   // a copy failed: the remaining pieces are never handed out
        std::lock_guard<std::mutex> lk(mu);
        j->pending -= j->n_pieces - handled;
      }
      // the raw head crosses while the host threads are busy with the packed pieces
      if (ok && j->n_raw > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        cudaError_t e = cudaMemcpyAsync(j->frames_host, j->frames_dev, (size_t)j->n_raw * frame_bytes, cudaMemcpyDeviceToHost,
                                        copy_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(copy_stream);
        if (e != cudaSuccess) { ok = false; err = "raw frame copy failed"; }
        std::lock_guard<std::mutex> lk(mu);
        j->s_raw = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        bytes_raw += (size_t)j->n_raw * frame_bytes;
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (any_packed) ++jobs_packed; else ++jobs_raw;
        job_part_done(j, ok, err);
      }
    }
  }

  void worker_loop() {
    long long my_last = -1;   // id of the last piece this thread worked on: every thread takes part in every piece once
    for (;;) {
      Work *w = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || (!q_work.empty() && q_work.front()->id > my_last); });
        if (q_work.empty() || q_work.front()->id <= my_last) return;
        w = q_work.front();
        my_last = w->id;
        if (!w->started) {
          w->started = true;
          w->t_start = std::chrono::steady_clock::now();
        }
      }
      const Job *j = w->job;
      const int c = w->piece;
      // frames are handed out in blocks so that a slow thread does not hold the piece back
      const long long blockf = 256;
      const long long p0 = j->piece_f0[c], p1 = j->piece_f0[c + 1];
      uint8_t *dst = j->frames_host + (size_t)j->n_raw * j->npx * 4;
      bool ok = true;
      for (;;) {
        const long long f0 = p0 + w->next.fetch_add(blockf);
        if (f0 >= p1) break;
        const long long f1 = f0 + blockf < p1 ? f0 + blockf : p1;
        ok &= unpack_range(h_tokens[j->slot] + j->piece_tok0[c], h_table[j->slot], f0, f1, j->npx, dst);
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) w->ok = false;
        if (--w->left == 0) {
          q_work.pop_front();
          Job *jj = w->job;
          jj->s_unp += std::chrono::duration<double>(std::chrono::steady_clock::now() - w->t_start).count();
          jj->frames_unpacked += p1 - p0;
          const bool good = w->ok;
          delete w;
          job_part_done(jj, good, "packed frame stream is inconsistent");
          cv_work.notify_all();
        }
      }
    }
  }
};

PackPipe *packpipe_create(int device) {
  PackPipe *p = new PackPipe;
  p->device = device;
  // host threads: the CPUs this process may run on, shared with the other ranks of the node when the process was
  // started by torchrun (one rank per GPU, all on the same host cores), at most 16
  unsigned hw = std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) {
    const unsigned allowed = (unsigned)CPU_COUNT(&set);
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
      const int ranks = atoi(e);
      if (ranks > 1 && allowed == hw) hw = allowed / (unsigned)ranks;   // not pinned: an equal share
      else hw = allowed;
    } else {
      hw = allowed;
    }
  }
  if (hw < 2) hw = 2;
  p->n_workers = hw > 16 ? 16 : hw;
  if (const char *e = getenv("PPB_UNPACK_THREADS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 256) p->n_workers = (unsigned)v;
  }
  if (const char *e = getenv("PPB_RAW_FRACTION")) {   // developer switch: a fixed split instead of the tracked one
    const double v = atof(e);
    if (v >= 0.0 && v <= 1.0) { p->raw_fraction = v; p->split_fixed = true; }
  }
  cudaSetDevice(device);
  if (cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete p; return nullptr; }
  for (int i = 0; i < 2; ++i)
    for (int c = 0; c < kMaxPieces; ++c)
      if (cudaEventCreateWithFlags(&p->ev_packed[i][c], cudaEventDisableTiming) != cudaSuccess) { delete p; return nullptr; }
  p->coordinator = std::thread([p] { p->coordinator_loop(); });
  for (unsigned i = 0; i < p->n_workers; ++i)
    p->workers.emplace_back([p] {
      // the caller's own thread (it uploads the next step's tables and enqueues its kernels) and the coordinator go
      // first when every CPU is busy rebuilding frames
      if (nice(5) == -1) { /* not permitted: same priority */ }
      p->worker_loop();
    });
  return p;
}

void packpipe_destroy(PackPipe *p) {
  if (!p) return;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return p->finished == p->submitted; });
    p->stop = true;
  }
  p->cv_coord.notify_all();
  p->cv_work.notify_all();
  p->cv_done.notify_all();
  if (p->coordinator.joinable()) p->coordinator.join();
  for (auto &t : p->workers) t.join();
  cudaSetDevice(p->device);
  for (int i = 0; i < 2; ++i) {
    if (p->d_tokens[i]) cudaFree(p->d_tokens[i]);
    if (p->d_table[i]) cudaFree(p->d_table[i]);
    if (p->d_ctrl[i]) cudaFree(p->d_ctrl[i]);
    if (p->h_tokens[i]) cudaFreeHost(p->h_tokens[i]);
    if (p->h_table[i]) cudaFreeHost(p->h_table[i]);
    if (p->h_ctrl[i]) cudaFreeHost(p->h_ctrl[i]);
    for (int c = 0; c < kMaxPieces; ++c)
      if (p->ev_packed[i][c]) cudaEventDestroy(p->ev_packed[i][c]);
  }
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  delete p;
}

// token capacity of a piece of nf frames: packed frames may take a quarter of the raw size at most
static size_t piece_capacity(long long nf, int npx) { return (size_t)nf * npx / 4 + 64; }

static cudaError_t reserve_slot(PackPipe *p, int s, long long n_frames, int npx) {
  const size_t want_tokens = piece_capacity(n_frames, npx) + 64 * kMaxPieces;
  cudaError_t e = cudaSuccess;
  if (want_tokens > p->cap_tokens[s]) {
    if (p->d_tokens[s]) cudaFree(p->d_tokens[s]);
    if (p->h_tokens[s]) cudaFreeHost(p->h_tokens[s]);
    p->d_tokens[s] = nullptr; p->h_tokens[s] = nullptr; p->cap_tokens[s] = 0;
    if ((e = cudaMalloc((void **)&p->d_tokens[s], want_tokens * 4)) != cudaSuccess) return e;
    if ((e = cudaHostAlloc((void **)&p->h_tokens[s], want_tokens * 4, cudaHostAllocDefault)) != cudaSuccess) return e;
    p->cap_tokens[s] = want_tokens;
  }
  if ((size_t)n_frames > p->cap_frames[s]) {
    if (p->d_table[s]) cudaFree(p->d_table[s]);
    if (p->h_table[s]) cudaFreeHost(p->h_table[s]);
    p->d_table[s] = nullptr; p->h_table[s] = nullptr; p->cap_frames[s] = 0;
    if ((e = cudaMalloc((void **)&p->d_table[s], (size_t)n_frames * sizeof(uint2))) != cudaSuccess) return e;
    if ((e = cudaHostAlloc((void **)&p->h_table[s], (size_t)n_frames * sizeof(uint2), cudaHostAllocDefault)) != cudaSuccess) return e;
    p->cap_frames[s] = (size_t)n_frames;
  }
  if (!p->d_ctrl[s]) {
    if ((e = cudaMalloc((void **)&p->d_ctrl[s], 16 * kMaxPieces)) != cudaSuccess) return e;
    if ((e = cudaHostAlloc((void **)&p->h_ctrl[s], 16 * kMaxPieces, cudaHostAllocDefault)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int packpipe_submit(PackPipe *p, int slot, const uint8_t *frames_dev, long long n_frames, int npx, uint8_t *frames_host,
                    cudaStream_t producer, std::string *err) {
  if (n_frames <= 0) return 0;
  double f;
  {
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv_done.wait(lk, [&] { return !p->slot_busy[slot]; });   // the slot's previous job has left the host buffers
    if (!p->error.empty()) { *err = p->error; return -1; }
    f = p->raw_fraction;
  }
  Job *j = new Job;
  j->slot = slot; j->frames_dev = frames_dev; j->frames_host = frames_host; j->n_frames = n_frames; j->npx = npx;
  // the head of the job goes raw through the copy engine, the rest is packed in pieces (small jobs: one packed piece)
  if (n_frames >= 4096) {
    j->n_raw = (long long)(f * (double)n_frames) / 256 * 256;
    if (j->n_raw > n_frames - 256) j->n_raw = n_frames - 256;
  }
  const long long n_packed = n_frames - j->n_raw;
  j->n_pieces = n_packed >= 8192 ? kMaxPieces : 1;
  for (int c = 0; c <= j->n_pieces; ++c) j->piece_f0[c] = c == j->n_pieces ? n_packed : (n_packed / j->n_pieces) / 256 * 256 * c;
  for (int c = 0; c < j->n_pieces; ++c)
    j->piece_tok0[c + 1] = j->piece_tok0[c] + piece_capacity(j->piece_f0[c + 1] - j->piece_f0[c], npx);
  j->pending = j->n_pieces + 1;
  const int n_launched = j->n_pieces;
  cudaError_t e = reserve_slot(p, slot, n_frames, npx);   // sized for the whole job: the split moves
  if (e == cudaSuccess) e = cudaMemsetAsync(p->d_ctrl[slot], 0, 16 * kMaxPieces, producer);
  for (int c = 0; c < j->n_pieces && e == cudaSuccess; ++c) {
    const long long f0 = j->piece_f0[c], nf = j->piece_f0[c + 1] - f0;
    long long blocks = (nf + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_frames_kernel<<<(unsigned)blocks, 256, 0, producer>>>(
        reinterpret_cast<const uint32_t *>(frames_dev + (size_t)(j->n_raw + f0) * npx * 4), nf, npx,
        p->d_tokens[slot] + j->piece_tok0[c], (uint32_t)(piece_capacity(nf, npx) - 64), p->d_table[slot] + f0,
        p->d_ctrl[slot] + 4 * c);
    e = cudaGetLastError();
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(p->h_ctrl[slot] + 4 * c, p->d_ctrl[slot] + 4 * c, 16, cudaMemcpyDeviceToHost, producer);
    if (e == cudaSuccess) e = cudaEventRecord(p->ev_packed[slot][c], producer);
  }
  if (e != cudaSuccess) { *err = cudaGetErrorString(e); delete j; return -1; }
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->slot_busy[slot] = true;
    ++p->submitted;
    p->q_coord.push_back(j);
  }
  p->cv_coord.notify_one();
  return n_launched;
}

long long packpipe_submitted(PackPipe *p) {
  std::lock_guard<std::mutex> lk(p->mu);
  return p->submitted;
}

// blocks until the first `target` jobs are done (jobs finish in submission order)
int packpipe_wait_finished(PackPipe *p, long long target, std::string *err) {
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv_done.wait(lk, [&] { return p->finished >= target; });
  if (!p->error.empty()) { *err = p->error; p->error.clear(); return -1; }
  return 0;
}

int packpipe_wait_slot(PackPipe *p, int slot, std::string *err) {
  std::unique_lock<std::mutex> lk(p->mu);
  p->cv_done.wait(lk, [&] { return !p->slot_busy[slot]; });
  if (!p->error.empty()) { *err = p->error; p->error.clear(); return -1; }
  return 0;
}

void packpipe_stats(PackPipe *p, long long out[6]) {
  std::lock_guard<std::mutex> lk(p->mu);
  out[0] = p->jobs_packed; out[1] = p->jobs_raw; out[2] = (long long)p->bytes_packed; out[3] = (long long)p->bytes_raw;
  out[4] = (long long)p->us_copy; out[5] = (long long)p->us_unpack;
}

}  // namespace ppb
