// Developer probe: how fast can T host threads fill a 2 GiB frame array (the caller's RGBA buffer of one bench step)
// with regular and with non-temporal stores?  Decides whether expanding packed frames on the host can beat the
// 51 GB/s a pinned D2H copy reaches.   g++ -O3 -pthread tools/host_fill_probe.cpp -o tools/_bin/host_fill_probe
#include <emmintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static void fill_regular(uint8_t *p, size_t n) { memset(p, 0xff, n); }
static void fill_nt(uint8_t *p, size_t n) {
  const __m128i v = _mm_set1_epi32(-1);
  for (size_t i = 0; i < n; i += 64) {
    _mm_stream_si128((__m128i *)(p + i), v);
    _mm_stream_si128((__m128i *)(p + i + 16), v);
    _mm_stream_si128((__m128i *)(p + i + 32), v);
    _mm_stream_si128((__m128i *)(p + i + 48), v);
  }
  _mm_sfence();
}
int main() {
  const size_t bytes = (size_t)2 << 30;
  uint8_t *buf = (uint8_t *)aligned_alloc(4096, bytes);
  memset(buf, 1, bytes);
  for (int mode = 0; mode < 2; ++mode)
    for (int T : {1, 2, 4, 8, 16, 32}) {
      double best = 0;
      for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
          th.emplace_back([=] {
            const size_t per = bytes / T;
            (mode ? fill_nt : fill_regular)(buf + t * per, per);
          });
        for (auto &x : th) x.join();
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (bytes / s / 1e9 > best) best = bytes / s / 1e9;
      }
      printf("%s stores, %2d threads: %.1f GB/s\n", mode ? "non-temporal" : "regular", T, best);
    }
  return buf[12345] == 0;
}
