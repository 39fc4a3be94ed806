// Developer probe / CPU test helper: speed and correctness of the host half of the packed frame download
// (csrc/ppb_pack.cu: unpack_range) on synthetic billiards-like frames, without a GPU.
//   nvcc -std=c++17 -O3 -o tools/_bin/unpack_host_check tools/unpack_host_check.cu && tools/_bin/unpack_host_check [threads]
#define PPB_PACK_HOST_CHECK 1
#include "../pyphy_engine_b200/csrc/ppb_pack.cu"
#include <chrono>
#include <cstdio>
#include <random>
#include <string>

int main(int argc, char **argv) {
  const int threads = argc > 1 ? atoi(argv[1]) : 8;
  const int npx = 4096;
  const long long n_frames = 65536;
  std::mt19937 rng(1);
  // a frame: white, a 3-pixel red border of a box, eight 8x8 "sprites" with 6 distinct values per row
  std::vector<uint32_t> frames((size_t)n_frames * npx, 0xffffffffu);
  for (long long f = 0; f < n_frames; ++f) {
    uint32_t *im = &frames[(size_t)f * npx];
    const int x0 = 3 + rng() % 6, y0 = 3 + rng() % 6, x1 = 55 + rng() % 6, y1 = 55 + rng() % 6;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x)
        if (y < y0 + 3 || y >= y1 - 3 || x < x0 + 3 || x >= x1 - 3) im[y * 64 + x] = 0xff0000ffu;
    for (int b = 0; b < 8; ++b) {
      const int bx = x0 + 4 + rng() % (x1 - x0 - 16), by = y0 + 4 + rng() % (y1 - y0 - 16);
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) {
          const int d = (x < 4 ? x : 7 - x) + (y < 4 ? y : 7 - y);
          if (d >= 2) im[(by + y) * 64 + bx + x] = d >= 4 ? 0xff808080u : (0xff000000u | (uint32_t)(40 * d + 7 * x + y) * 0x010101u);
        }
    }
  }
  // the encoder the kernel implements, sequentially
  std::vector<uint32_t> tokens;
  std::vector<uint2> table(n_frames);
  for (long long f = 0; f < n_frames; ++f) {
    const uint32_t *im = &frames[(size_t)f * npx];
    const uint32_t off = (uint32_t)tokens.size();
    for (int base = 0; base < npx; base += 64) {
      int p = 0;
      const int valid = std::min(64, npx - base);
      while (p < valid) {
        int q = p + 1;
        while (q < valid && im[base + q] == im[base + p]) ++q;
        tokens.push_back((im[base + p] & 0xffffffu) | ((uint32_t)(q - p) << 24));
        p = q;
      }
    }
    table[f] = make_uint2(off, (uint32_t)tokens.size() - off);
  }
  printf("%.0f tokens per frame, %.2f KB per frame\n", (double)tokens.size() / n_frames, 4.0 * tokens.size() / n_frames / 1024);
  // destination: ordinary memory, or (second argument "pinned", needs a CUDA device) page-locked memory like the
  // caller's array of ppb_simulate_host
  const bool pinned = argc > 2 && std::string(argv[2]) == "pinned";
  uint8_t *out = nullptr;
  if (pinned) {
    if (cudaHostAlloc((void **)&out, (size_t)n_frames * npx * 4, cudaHostAllocDefault) != cudaSuccess) {
      printf("cudaHostAlloc failed (no device?)\n");
      return 2;
    }
  } else {
    out = (uint8_t *)aligned_alloc(4096, (size_t)n_frames * npx * 4);
  }
  memset(out, 0, (size_t)n_frames * npx * 4);
  double best = 1e9;
  for (int rep = 0; rep < 4; ++rep) {
    // the threads exist before the clock starts (the library's pool is persistent); they leave a spin barrier together
    std::vector<std::thread> th;
    std::atomic<int> bad{0}, ready{0};
    std::atomic<bool> go{false};
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&, t] {
        ready++;
        while (!go.load(std::memory_order_acquire)) {}
        const long long per = n_frames / threads;
        if (!ppb::host_check_unpack(tokens.data(), table.data(), t * per, t == threads - 1 ? n_frames : (t + 1) * per, npx, out)) bad++;
      });
    while (ready.load() < threads) {}
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    for (auto &x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (s < best) best = s;
    if (bad) { printf("inconsistent stream\n"); return 1; }
  }
  const bool same = memcmp(out, frames.data(), (size_t)n_frames * npx * 4) == 0;
  printf("%d threads%s: %.1f GB/s of RGBA rebuilt, %s\n", threads, pinned ? " (pinned destination)" : "",
         (double)n_frames * npx * 4 / best / 1e9, same ? "identical" : "DIFFERENT");
  return same ? 0 : 1;
}
