#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
// mode 0: xmm NT fill; 1: zmm NT fill; 2: zmm load from 16KB L1 block + NT store; 3: mode 2 + rewrite the block with regular stores first (decoder-like)
static void work(int mode, uint8_t *p, size_t n) {
  alignas(64) static thread_local uint32_t blk[2][4096 + 128];
  const __m512i v = _mm512_set1_epi32(-1);
  if (mode == 0) {
    const __m128i x = _mm_set1_epi32(-1);
    for (size_t i = 0; i < n; i += 64) { _mm_stream_si128((__m128i*)(p+i), x); _mm_stream_si128((__m128i*)(p+i+16), x); _mm_stream_si128((__m128i*)(p+i+32), x); _mm_stream_si128((__m128i*)(p+i+48), x);} 
  } else if (mode == 1) {
    for (size_t i = 0; i < n; i += 64) _mm512_stream_si512((__m512i*)(p+i), v);
  } else {
    int cur = 0;
    for (size_t f = 0; f < n; f += 16384) {
      if (mode == 3) { uint32_t *b = blk[cur]; for (int k = 0; k < 4096; k += 7) { _mm512_storeu_si512(b + k, v); _mm512_storeu_si512(b + k + 16, v);} }
      const uint32_t *s = blk[cur ^ (mode == 3)];
      for (int l = 0; l < 256; ++l) _mm512_stream_si512((__m512i*)(p + f + l*64), _mm512_load_si512(s + l*16));
      cur ^= 1;
    }
  }
  _mm_sfence();
}
int main() {
  const size_t bytes = (size_t)1 << 30;
  uint8_t *buf = (uint8_t*)aligned_alloc(4096, bytes); memset(buf, 1, bytes);
  for (int mode = 0; mode < 4; ++mode) for (int T : {1, 2, 4, 8}) {
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
      std::vector<std::thread> th; std::atomic<int> ready{0}; std::atomic<bool> go{false};
      for (int t = 0; t < T; ++t) th.emplace_back([&, t]{ ready++; while(!go.load()){} size_t per = bytes / T; work(mode, buf + t*per, per); });
      while (ready.load() < T) {}
      auto t0 = std::chrono::steady_clock::now(); go.store(true);
      for (auto &x : th) x.join();
      double s = std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
      if (bytes/s/1e9 > best) best = bytes/s/1e9;
    }
    printf("mode %d threads %d: %.1f GB/s\n", mode, T, best);
  }
}
