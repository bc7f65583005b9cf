// host_pack_probe.cpp -- how fast can the HOST turn ASCII reads into 2-bit codes?  (decides whether
// packing before the PCIe copy can beat sending the ASCII bytes: 1.5 GB at ~52 GB/s = 29 ms)
// g++ -O3 -march=native -pthread host_pack_probe.cpp -o host_pack_probe
#include <immintrin.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// 32 ASCII bytes -> 8 bytes of 2-bit codes ((c >> 1) & 3), plus a validity flag (all in ACGTacgt)
static inline bool pack32(const uint8_t* src, uint8_t* dst) {
  __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src));
  __m256i up = _mm256_and_si256(v, _mm256_set1_epi8(char(0xDF)));
  __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('C'))),
                               _mm256_or_si256(_mm256_cmpeq_epi8(up, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(up, _mm256_set1_epi8('T'))));
  // crumbs: bits 1..2 of every byte
  alignas(32) uint64_t w[4];
  _mm256_store_si256(reinterpret_cast<__m256i*>(w), v);
  uint64_t out = 0;
  for (int i = 0; i < 4; i++) out |= _pext_u64(w[i], 0x0606060606060606ull) << (16 * i);
  memcpy(dst, &out, 8);
  return _mm256_movemask_epi8(ok) == -1;
}

int main(int argc, char** argv) {
  size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1500000000ull;
  int max_threads = argc > 2 ? atoi(argv[2]) : int(std::thread::hardware_concurrency());
  uint8_t* src = static_cast<uint8_t*>(aligned_alloc(4096, n + 64));
  uint8_t* dst = static_cast<uint8_t*>(aligned_alloc(4096, n / 4 + 64));
  const char L[4] = {'A', 'C', 'G', 'T'};
  uint64_t x = 88172645463325252ull;
  for (size_t i = 0; i < n; i++) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    src[i] = uint8_t(L[x & 3]);
  }
  memset(dst, 0, n / 4 + 64);
  for (int nt : {1, 2, 4, 8, 12, 16, 24, 32}) {
    if (nt > max_threads) break;
    double best = 1e9;
    for (int it = 0; it < 4; it++) {
      double t0 = now();
      std::vector<std::thread> th;
      std::vector<int> bad(nt, 0);
      for (int t = 0; t < nt; t++)
        th.emplace_back([&, t] {
          size_t lo = (n / 32 * t / nt) * 32, hi = (n / 32 * (t + 1) / nt) * 32;
          bool ok = true;
          for (size_t i = lo; i < hi; i += 32) ok &= pack32(src + i, dst + i / 4);
          bad[t] = !ok;
        });
      for (auto& q : th) q.join();
      best = std::min(best, now() - t0);
    }
    printf("threads %2d: %.1f ms for %.2f GB ASCII = %.1f GB/s\n", nt, best * 1e3, n / 1e9, n / best / 1e9);
  }
  // memcpy ceiling for scale
  double t0 = now();
  memcpy(dst, src, n / 4);
  printf("single-thread memcpy of n/4: %.1f GB/s\n", n / 4 / (now() - t0) / 1e9);
  return 0;
}
