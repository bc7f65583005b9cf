// hostpack.cpp -- host-side 2-bit packing of nucleotide queries ahead of the PCIe copy.
//
// Why: parallel_count's input is ASCII (`&str`, /root/reference/src/fm_index.rs:455-460); 10 M x 150-bp
// reads are 1.5 GB, which a PCIe 5 x16 link moves in ~29 ms -- longer than the 20 ms the search kernel
// needs.  The host cores pack the bases to 2 bits while earlier chunks are in flight, so only a quarter
// of the bytes cross the link; the device expands them to its 4-bit search-order words (pack2_kernel).
// Plain C++ (g++), no CUDA: SIMD levels are selected at run time.
#include "hostpack.hpp"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>

namespace awry {

namespace {

// ---------------------------------------------------------------- thread pool
// A shared work-queue pool.  A parallel region is a JOB: `n_blocks` independent blocks handed out through an
// atomic counter.  Several jobs may be open at once -- in a multi-replica call every replica's host thread
// packs its own chunks (/root/reference/src/fm_index.rs:455-460 is ONE call over ONE batch, whatever the
// number of GPUs behind it) -- and every worker serves the oldest job that still has blocks, so the
// packing capacity of the whole host follows the demand instead of being split per caller.  The caller
// works on its own job too and returns when all its blocks are done.  The pool grows on demand up to the
// thread cap (host_set_threads); an exception thrown by a block is caught in the worker, remembered in
// the job and rethrown from run() in the caller's thread, so nothing unwinds out of a pool thread.
class Pool {
 public:
  struct Job {
    const std::function<void(size_t)>* fn = nullptr;
    size_t n_blocks = 0;
    std::atomic<size_t> next{0};
    std::atomic<size_t> done{0};
    std::atomic<bool> failed{false};
    std::exception_ptr error;
    std::mutex err_mu;
  };

  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int cap() {
    std::lock_guard<std::mutex> lk(mu_);
    return cap_;
  }
  void set_cap(int n) {
    std::lock_guard<std::mutex> lk(mu_);
    cap_ = std::max(1, std::min(n, 256));
  }
  // runs fn(block) for block in [0, n_blocks) on up to `max_threads` threads (0 = the pool's cap)
  // open_before: parallel regions already open when this one starts (its competitors for the pool's threads)
  void run(size_t n_blocks, const std::function<void(size_t)>& fn, int max_threads = 0, int* open_before = nullptr) {
    if (open_before) *open_before = 0;
    if (n_blocks == 0) return;
    auto job = std::make_shared<Job>();
    job->fn = &fn;
    job->n_blocks = n_blocks;
    int want;
    {
      std::lock_guard<std::mutex> lk(mu_);
      want = max_threads > 0 ? std::min(max_threads, cap_) : cap_;
      want = int(std::min<size_t>(size_t(want), n_blocks));
      // the caller is one of the threads; workers are shared between the open jobs
      while (int(workers_.size()) + 1 < want) {
        try {
          workers_.emplace_back([this] { loop(); });
        } catch (...) {
          break;  // cannot spawn: run with what there is
        }
      }
      if (open_before) *open_before = int(jobs_.size());
      if (want > 1) jobs_.push_back(job);
    }
    if (want > 1) cv_.notify_all();
    work(*job);
    if (want > 1) {
      std::unique_lock<std::mutex> lk(mu_);
      done_cv_.wait(lk, [&] { return job->done.load(std::memory_order_acquire) == job->n_blocks; });
      jobs_.erase(std::remove(jobs_.begin(), jobs_.end(), job), jobs_.end());
    }
    if (job->error) std::rethrow_exception(job->error);
  }

 private:
  // takes blocks of `job` until none are left; returns after the last one it ran
  void work(Job& job) {
    for (;;) {
      size_t b = job.next.fetch_add(1, std::memory_order_relaxed);
      if (b >= job.n_blocks) return;
      if (!job.failed.load(std::memory_order_relaxed)) {
        try {
          (*job.fn)(b);
        } catch (...) {
          std::lock_guard<std::mutex> lk(job.err_mu);
          if (!job.error) job.error = std::current_exception();
          job.failed.store(true, std::memory_order_relaxed);
        }
      }
      if (job.done.fetch_add(1, std::memory_order_acq_rel) + 1 == job.n_blocks) {
        std::lock_guard<std::mutex> lk(mu_);  // pairs with the waiter's predicate check
        done_cv_.notify_all();
      }
    }
  }
  void loop() {
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] {
          if (quit_) return true;
          for (auto& j : jobs_)
            if (j->next.load(std::memory_order_relaxed) < j->n_blocks) {
              job = j;
              return true;
            }
          return false;
        });
        if (quit_) return;
      }
      work(*job);
    }
  }
  std::vector<std::thread> workers_;
  std::vector<std::shared_ptr<Job>> jobs_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  int cap_ = 1;
  bool quit_ = false;
};

int default_threads() {
  if (const char* e = getenv("AWRY_B200_HOST_THREADS")) {
    int v = atoi(e);
    if (v >= 1) return std::min(v, 256);
  }
  int hw = int(std::max(1u, std::thread::hardware_concurrency()));
  int local = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));  // one process per GPU
  return std::max(1, std::min(16, hw / local));
}

Pool& pool() {
  static Pool* p = [] {
    auto* q = new Pool();  // never destroyed: worker threads must not be joined from a static destructor
    q->set_cap(default_threads());
    return q;
  }();
  return *p;
}

// ---------------------------------------------------------------- packers
inline bool is_acgt(uint8_t c) {
  c &= 0xDF;
  return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

void pack_scalar(const uint8_t* src, size_t lo, size_t hi, uint8_t* dst, std::vector<uint64_t>& exc) {
  for (size_t i = lo; i < hi; i++) {
    uint8_t c = src[i];
    if (!is_acgt(c)) exc.push_back((uint64_t(i) << 8) | c);
    uint8_t crumb = uint8_t((c >> 1) & 3u);
    uint8_t sh = uint8_t(2 * (i & 3));
    dst[i >> 2] = uint8_t((dst[i >> 2] & ~(3u << sh)) | (crumb << sh));
  }
}

// [lo, hi) with lo % 32 == 0 and (hi - lo) % 32 == 0.
// Main loop: 128 bases -> 32 bytes per iteration, all in 256-bit SIMD: crumbs = (c >> 1) & 3 per byte, two
// multiply-adds fold 4 crumbs into one byte-sized value per 32-bit lane (pmaddubsw x (1,4), pmaddwd x (1,16)),
// two saturating packs and one cross-lane permute put the 32 bytes in order.  Validity (is every byte one of
// ACGTacgt?) is one table lookup per byte: pshufb indexes a 16-entry table with the low nibble of the
// upper-cased byte ('A' 0x41 -> 1, 'C' 0x43 -> 3, 'T' 0x54 -> 4, 'G' 0x47 -> 7; bit 7 set -> 0) and the result
// must equal the byte.  ~42 uops per 128 bases against ~100 for the PEXT formulation this replaces
// (4 x vextract + 4 x PEXT per 32 bases), plus a software prefetch 2 KiB ahead (the hardware streamer stops at
// page boundaries).  One core of the development host: 8.5 -> 18 GB/s with the input in L3, 4.2 -> 9.4 GB/s
// from DRAM (6.3 without the prefetch).  The 32-base PEXT step remains for the tail of a block.
__attribute__((target("avx2,bmi2"))) void pack_avx2(const uint8_t* src, size_t lo, size_t hi, uint8_t* dst,
                                                      std::vector<uint64_t>& exc) {
  const __m256i up_mask = _mm256_set1_epi8(char(0xDF));
  // unused entries hold 0x20: an upper-cased byte never has bit 5 set, so e.g. NUL or ' ' (-> 0x00) cannot match
  constexpr char X = 0x20;
  const __m256i lut = _mm256_setr_epi8(X, 'A', X, 'C', 'T', X, X, 'G', X, X, X, X, X, X, X, X,
                                       X, 'A', X, 'C', 'T', X, X, 'G', X, X, X, X, X, X, X, X);
  const __m256i three = _mm256_set1_epi8(3);
  const __m256i mul8 = _mm256_set1_epi16(0x0401);       // b0 + 4 * b1   (pmaddubsw: unsigned a * signed b)
  const __m256i mul16 = _mm256_set1_epi32(0x00100001);  // w0 + 16 * w1  (pmaddwd)
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  constexpr size_t PF = 2048;  // software prefetch distance (bytes)
  // The packed bytes are read next by the PCIe DMA engine, not by a core: streaming stores keep them out of
  // the caches and spare the read-for-ownership of every output line -- with several GPUs behind one host the
  // packer is bound by host memory traffic (1 B read + 0.25 B written + 0.25 B read back by the DMA per base;
  // a regular store adds another 0.25 B of ownership reads).  Needs a 32-byte aligned destination.
  const bool stream = (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
  size_t i = lo;
  for (; i + 128 <= hi; i += 128) {
    __m256i v[4], d[4], ok = _mm256_set1_epi8(char(0xFF));
    _mm_prefetch(reinterpret_cast<const char*>(src + i + PF), _MM_HINT_T0);
    _mm_prefetch(reinterpret_cast<const char*>(src + i + PF + 64), _MM_HINT_T0);
#pragma GCC unroll 4
    for (int j = 0; j < 4; j++) {
      v[j] = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32 * j));
      const __m256i up = _mm256_and_si256(v[j], up_mask);
      ok = _mm256_and_si256(ok, _mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, up), up));
      const __m256i crumbs = _mm256_and_si256(_mm256_srli_epi16(v[j], 1), three);
      d[j] = _mm256_madd_epi16(_mm256_maddubs_epi16(crumbs, mul8), mul16);  // 8 x u32, each one output byte
    }
    // per 128-bit lane: [d0 d1 d2 d3] of the low / high four dwords -> bytes; then interleave the two lanes
    const __m256i q = _mm256_packus_epi16(_mm256_packus_epi32(d[0], d[1]), _mm256_packus_epi32(d[2], d[3]));
    if (stream)
      _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), _mm256_permutevar8x32_epi32(q, order));
    else
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + (i >> 2)), _mm256_permutevar8x32_epi32(q, order));
    if (__builtin_expect(_mm256_movemask_epi8(ok) != -1, 0)) {  // rare: list the bytes outside ACGTacgt
#pragma GCC unroll 4
      for (int j = 0; j < 4; j++) {
        const __m256i up = _mm256_and_si256(v[j], up_mask);
        uint32_t bad = ~uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, up), up)));
        while (bad) {
          const unsigned b = unsigned(__builtin_ctz(bad));
          bad &= bad - 1;
          exc.push_back((uint64_t(i + 32 * j + b) << 8) | src[i + 32 * j + b]);
        }
      }
    }
  }
  if (stream) _mm_sfence();  // the streaming stores are globally visible before the block is reported done
  for (; i < hi; i += 32) {  // tail of the block: one 32-base step at a time
    __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    __m256i up = _mm256_and_si256(v, up_mask);
    uint32_t bad = ~uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, up), up)));
    while (bad) {
      unsigned b = unsigned(__builtin_ctz(bad));
      bad &= bad - 1;
      exc.push_back((uint64_t(i + b) << 8) | src[i + b]);
    }
    uint64_t w0 = uint64_t(_mm256_extract_epi64(v, 0)), w1 = uint64_t(_mm256_extract_epi64(v, 1)),
             w2 = uint64_t(_mm256_extract_epi64(v, 2)), w3 = uint64_t(_mm256_extract_epi64(v, 3));
    const uint64_t M = 0x0606060606060606ull;
    uint64_t out = _pext_u64(w0, M) | (_pext_u64(w1, M) << 16) | (_pext_u64(w2, M) << 32) | (_pext_u64(w3, M) << 48);
    memcpy(dst + (i >> 2), &out, 8);
  }
}

// [lo, hi) with lo % 64 == 0 and (hi - lo) % 64 == 0
__attribute__((target("avx512f,avx512bw"))) void pack_avx512(const uint8_t* src, size_t lo, size_t hi, uint8_t* dst,
                                                               std::vector<uint64_t>& exc) {
  const __m512i up_mask = _mm512_set1_epi8(char(0xDF));
  const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'),
                cT = _mm512_set1_epi8('T');
  const __m512i three = _mm512_set1_epi8(3);
  const __m512i mul8 = _mm512_set1_epi16(0x0401);   // b0 + 4*b1      (pmaddubsw: unsigned a * signed b)
  const __m512i mul16 = _mm512_set1_epi32(0x00100001);  // w0 + 16*w1 (pmaddwd)
  for (size_t i = lo; i < hi; i += 64) {
    __m512i v = _mm512_loadu_si512(src + i);
    __m512i up = _mm512_and_si512(v, up_mask);
    __mmask64 ok = _mm512_cmpeq_epi8_mask(up, cA) | _mm512_cmpeq_epi8_mask(up, cC) | _mm512_cmpeq_epi8_mask(up, cG) |
                   _mm512_cmpeq_epi8_mask(up, cT);
    uint64_t bad = ~uint64_t(ok);
    while (bad) {
      unsigned b = unsigned(__builtin_ctzll(bad));
      bad &= bad - 1;
      exc.push_back((uint64_t(i + b) << 8) | src[i + b]);
    }
    __m512i crumbs = _mm512_and_si512(_mm512_srli_epi16(v, 1), three);  // per byte: (c >> 1) & 3
    __m512i w = _mm512_maddubs_epi16(crumbs, mul8);                      // 16-bit: c0 | c1 << 2
    __m512i d = _mm512_madd_epi16(w, mul16);                             // 32-bit: c0 | c1<<2 | c2<<4 | c3<<6
    __m128i bytes = _mm512_cvtepi32_epi8(d);                             // 16 packed bytes
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (i >> 2)), bytes);
  }
}

enum Level { SCALAR = 0, AVX2 = 1, AVX512 = 2 };
Level simd_level() {
  static Level l = [] {
    if (const char* e = getenv("AWRY_B200_HOST_SIMD")) {  // tests: 0 scalar, 1 avx2, 2 avx512
      int v = atoi(e);
      if (v == 0) return SCALAR;
      if (v == 1 && __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) return AVX2;
      if (v == 2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return AVX512;
    }
    // AVX2 first: measured 121 GB/s vs 95 GB/s for the AVX-512BW path on the 16-thread Sapphire Rapids host
    // of the B200 box (profiles/r01_s54_host_pack_rate_new_packer.log; 74 vs 63 GB/s with the PEXT packer)
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) return AVX2;
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw")) return AVX512;
    return SCALAR;
  }();
  return l;
}

}  // namespace

bool host_pack_supported() { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2"); }
int host_pool_threads() { return pool().cap(); }
void host_set_threads(int n) { pool().set_cap(n <= 0 ? default_threads() : n); }
int host_default_threads() { return default_threads(); }
void host_parallel_blocks(size_t n_blocks, const std::function<void(size_t)>& fn, int max_threads) {
  pool().run(n_blocks, fn, max_threads);
}

bool host_pack_dna(const uint8_t* src, size_t n, uint8_t* dst, std::vector<uint64_t>& exceptions, size_t max_exc_div,
                   int max_threads, int* sharers) {
  exceptions.clear();
  if (sharers) *sharers = 1;
  if (n == 0) return true;
  const Level lvl = simd_level();
  const size_t G = 64;  // granule: both SIMD widths and whole output bytes
  const size_t n_gran = n / G;
  // blocks of 256 KiB handed out dynamically: on a shared host the threads do not run at the same speed
  const size_t BLOCK = (256u << 10) / G;  // granules per block
  const size_t n_blocks = (n_gran + BLOCK - 1) / BLOCK;
  // exceptions are collected per block (blocks are independent jobs of the shared pool); once their number
  // passes the cut-off the chunk will travel as ASCII anyway, so the lists stop growing
  const size_t exc_limit = max_exc_div ? n / max_exc_div + 16 : ~size_t(0);
  std::vector<std::vector<uint64_t>> exc(n_blocks + 1);
  std::atomic<size_t> n_exc{0};
  auto body = [&](size_t b) {
    if (n_exc.load(std::memory_order_relaxed) > exc_limit) return;
    const size_t g0 = b * BLOCK;
    const size_t lo = g0 * G, hi = std::min(n_gran, g0 + BLOCK) * G;
    std::vector<uint64_t>& e = exc[b];
    if (lvl == AVX2)
      pack_avx2(src, lo, hi, dst, e);
    else if (lvl == AVX512)
      pack_avx512(src, lo, hi, dst, e);
    else
      pack_scalar(src, lo, hi, dst, e);
    if (!e.empty()) n_exc.fetch_add(e.size(), std::memory_order_relaxed);
  };
  int open_before = 0;
  pool().run(n_blocks, body, max_threads, &open_before);
  if (sharers) *sharers = open_before + 1;
  if (n_gran * G < n) {  // tail
    memset(dst + (n_gran * G) / 4, 0, (n - n_gran * G + 3) / 4);
    pack_scalar(src, n_gran * G, n, dst, exc[n_blocks]);
    n_exc.fetch_add(exc[n_blocks].size(), std::memory_order_relaxed);
  }
  const size_t total = n_exc.load();
  if (total > exc_limit) return false;
  exceptions.reserve(total);
  for (auto& e : exc) exceptions.insert(exceptions.end(), e.begin(), e.end());  // ascending by position
  return true;
}

}  // namespace awry
