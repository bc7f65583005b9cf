// api.cu -- host side of libawry_b200: `.awry` v1 loader, device replicas, the batched
// count / locate pipelines and the extern "C" boundary declared in include/awry_b200.h.
//
// Reference host code this stands in for (paths under /root/reference/src):
//   FmIndex::load / read_fm_index_by_version_number   fm_index_file.rs:132-160, :184-287
//   KmerLookupTable::from_file (skipped, Q1/Q2)        kmer_lookup_table.rs:55-77
//   SequenceIndex::from_file                           sequence_index.rs:155-183
//   parallel_count / parallel_locate (rayon map)       fm_index.rs:455-487
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <condition_variable>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <vector>

#include "../../include/awry_b200.h"
#include "build.hpp"
#include "hostpack.hpp"
#include "kernels.hpp"
#include "reads.hpp"

using namespace awry;

namespace {

thread_local char g_err[1024] = "";

struct ApiError : std::runtime_error {
  int code;
  ApiError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[900];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw ApiError(code, buf);
}

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      fail(AWRY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

template <class F>
int guarded(F&& f) {
  try {
    f();
    return AWRY_OK;
  } catch (const ApiError& e) {
    snprintf(g_err, sizeof g_err, "%s", e.what());
    return e.code;
  } catch (const std::bad_alloc&) {
    snprintf(g_err, sizeof g_err, "out of host memory");
    return AWRY_ERR_NOMEM;
  } catch (const std::exception& e) {
    snprintf(g_err, sizeof g_err, "%s", e.what());
    return AWRY_ERR_INVALID_ARG;
  } catch (...) {
    snprintf(g_err, sizeof g_err, "unknown error");
    return AWRY_ERR_INVALID_ARG;
  }
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    CU(cudaSetDevice(dev));
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ------------------------------------------------------------------ profiling

struct ProfState {
  std::mutex mu;
  bool enabled = false;
  struct Span {
    cudaEvent_t a, b;
    int kind;  // 0 search, 1 walk, 2 pack
    int device;
  };
  std::vector<Span> spans;
  uint64_t n[3] = {0, 0, 0};
  double ms[3] = {0, 0, 0};
  std::atomic<uint64_t> h2d{0}, d2h{0};
} g_prof;

struct ProfScope {
  bool on;
  cudaStream_t st;
  ProfState::Span sp{};
  ProfScope(int kind, int device, cudaStream_t s) : st(s) {
    on = g_prof.enabled;
    if (!on) return;
    sp.kind = kind;
    sp.device = device;
    cudaEventCreate(&sp.a);
    cudaEventCreate(&sp.b);
    cudaEventRecord(sp.a, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(sp.b, st);
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.spans.push_back(sp);
  }
};

void prof_collect_locked() {
  int prev = -1;
  cudaGetDevice(&prev);
  for (auto& s : g_prof.spans) {
    cudaSetDevice(s.device);
    float ms = 0;
    if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
      g_prof.ms[s.kind] += ms;
      g_prof.n[s.kind]++;
    }
    cudaEventDestroy(s.a);
    cudaEventDestroy(s.b);
  }
  g_prof.spans.clear();
  if (prev >= 0) cudaSetDevice(prev);
}

SearchVariant g_variant;  // experiments only (awry_set_search_variant)
int g_host_pack = -1;       // -1 = auto (AWRY_B200_HOST_PACK, CPU support, >= 4 pool threads), 0 = off, 1 = on
int g_locate_variant = 0;   // 0 = unsampled-SA gather when the array exists, 1 = always LF-walk (awry_set_locate_variant)


// ------------------------------------------------------------------ index

struct Workspace {
  int device = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t done = nullptr;
  // pinned staging
  uint8_t* h_qbytes = nullptr;
  size_t h_qbytes_cap = 0;
  uint64_t* h_qoff = nullptr;
  size_t h_qoff_cap = 0;  // entries
  uint8_t* h_out = nullptr;
  size_t h_out_cap = 0;  // bytes
  unsigned long long* h_flag = nullptr;
  unsigned long long* h_total = nullptr;  // hit total of a locate chunk (second word of the h_flag allocation)
  uint64_t* h_exc = nullptr;  // host-packed chunks: exception list (bytes outside ACGT)
  size_t h_exc_cap = 0;
  std::vector<uint64_t> exc_tmp;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;  // timing of a raw chunk's H2D copy (PackBalance)
  uint64_t link_probe_bytes = 0;
  // device
  uint8_t* d_qbytes = nullptr;
  size_t d_qbytes_cap = 0;
  uint64_t* d_qoff = nullptr;
  size_t d_qoff_cap = 0;
  uint64_t* d_qwords = nullptr;
  size_t d_qwords_cap = 0;
  uint32_t* d_defer = nullptr;
  size_t d_defer_cap = 0;
  uint8_t* d_out = nullptr;
  size_t d_out_cap = 0;
  uint64_t* d_hit_off = nullptr;
  size_t d_hit_off_cap = 0;
  void* d_temp = nullptr;
  size_t d_temp_cap = 0;
  unsigned long long* d_flag = nullptr;
  uint64_t* d_exc = nullptr;
  size_t d_exc_cap = 0;
  // reads-file front-end scratch, kept across calls (pinned allocations cost ~0.4 ms per MiB)
  struct ReadsScratch {
    uint64_t chunk = 0, carry = 0;
    uint8_t* h_buf[3] = {nullptr, nullptr, nullptr};
    uint8_t *d_raw = nullptr, *d_qbytes = nullptr;
    uint32_t *d_nl = nullptr, *d_small = nullptr, *d_seq_len = nullptr, *d_is_hdr = nullptr, *d_hdr_rank = nullptr;
    uint64_t *d_seq_off = nullptr, *d_qoff = nullptr;
    size_t cap_lines = 0, cap_qoff = 0, temp_bytes = 0;
    void* d_temp = nullptr;
    void* h_plan = nullptr;
    void release() {
      for (auto& b : h_buf) {
        cudaFreeHost(b);
        b = nullptr;
      }
      cudaFree(d_raw);
      cudaFree(d_qbytes);
      cudaFree(d_nl);
      cudaFree(d_small);
      cudaFree(d_seq_len);
      cudaFree(d_is_hdr);
      cudaFree(d_hdr_rank);
      cudaFree(d_seq_off);
      cudaFree(d_qoff);
      cudaFree(d_temp);
      cudaFreeHost(h_plan);
      *this = ReadsScratch();
    }
  } rs;

  template <class T>
  static void grow_dev(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = need + need / 8 + 256;
    CU(cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T)));
    cap = want;
  }
  template <class T>
  static void grow_host(T*& p, size_t& cap, size_t need) {
    if (need <= cap) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = need + need / 8 + 256;
    CU(cudaHostAlloc(reinterpret_cast<void**>(&p), want * sizeof(T), cudaHostAllocDefault));
    cap = want;
  }
  void init(int dev) {
    device = dev;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    CU(cudaHostAlloc(reinterpret_cast<void**>(&h_flag), 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    h_total = h_flag + 1;
    CU(cudaMalloc(reinterpret_cast<void**>(&d_flag), sizeof(unsigned long long)));
  }
  void destroy() {
    cudaSetDevice(device);
    if (st) cudaStreamSynchronize(st);
    cudaFreeHost(h_qbytes);
    cudaFreeHost(h_qoff);
    cudaFreeHost(h_out);
    cudaFreeHost(h_flag);
    cudaFreeHost(h_exc);
    cudaFree(d_exc);
    cudaFree(d_qbytes);
    cudaFree(d_qoff);
    cudaFree(d_qwords);
    cudaFree(d_defer);
    cudaFree(d_out);
    cudaFree(d_hit_off);
    cudaFree(d_temp);
    cudaFree(d_flag);
    rs.release();
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_b) cudaEventDestroy(ev_b);
    if (done) cudaEventDestroy(done);
    if (st) cudaStreamDestroy(st);
  }
};

struct Replica {
  int device = 0;
  int sm_count = 148;
  uint4* d_blocks = nullptr;
  uint64_t* d_sa = nullptr;
  uint2* d_table = nullptr;
  uint4* d_pair = nullptr;
  uint32_t* d_full_sa = nullptr;  // unsampled suffix array (locate accelerator)
  uint64_t* d_seq_starts = nullptr;
  unsigned long long* d_async_flag = nullptr;  // first bad query seen by *_device calls
  size_t bytes_blocks = 0, bytes_sa = 0, bytes_table = 0, bytes_pair = 0, bytes_full_sa = 0;
  uint32_t c2[16] = {0};
  IndexView view{};
  std::mutex ws_mu;
  std::vector<Workspace*> free_ws;
  std::vector<Workspace*> all_ws;

  Workspace* acquire() {
    {
      std::lock_guard<std::mutex> lk(ws_mu);
      if (!free_ws.empty()) {
        Workspace* w = free_ws.back();
        free_ws.pop_back();
        return w;
      }
    }
    auto* w = new Workspace();
    w->init(device);
    std::lock_guard<std::mutex> lk(ws_mu);
    all_ws.push_back(w);
    return w;
  }
  void release(Workspace* w) {
    std::lock_guard<std::mutex> lk(ws_mu);
    free_ws.push_back(w);
  }
};

}  // namespace

struct awry_index {
  uint64_t version = 1, sa_ratio = 0, bwt_len = 0;
  int alphabet = 0;
  int card = 6;
  uint32_t kmer_len_file = 0, kmer_len_dev = 0;
  uint64_t prefix_sums[23] = {0};
  uint64_t n_sa_words = 0;
  uint32_t sa_bits = 0;
  std::vector<uint64_t> seq_starts;
  std::vector<std::string> headers;
  std::vector<std::unique_ptr<Replica>> reps;
};

namespace {

uint32_t bits_per_element(uint64_t bwt_len) {  // compressed_suffix_array.rs:124-130
  uint64_t v = bwt_len - 1;
  return v ? 64u - uint32_t(__builtin_clzll(v)) : 0u;
}
uint64_t sa_word_len(uint64_t bwt_len, uint64_t ratio) {  // compressed_suffix_array.rs:113-123
  unsigned __int128 t = (unsigned __int128)((bwt_len + ratio - 1) / ratio) * bits_per_element(bwt_len);
  return uint64_t((t + 63) / 64);
}
uint64_t ipow(uint64_t b, uint32_t e) {
  uint64_t r = 1;
  while (e--) r *= b;
  return r;
}

// threads of a parallel positional read (AWRY_B200_IO_THREADS, default min(8, cores))
unsigned io_threads() {
  static const unsigned v = [] {
    if (const char* e = getenv("AWRY_B200_IO_THREADS")) return unsigned(std::min(64l, std::max(1l, strtol(e, nullptr, 10))));
    return std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  }();
  return v;
}

// sequential byte source: a file (FmIndex::load) or caller memory sections (FmIndex::new hand-over)
struct Source {
  FILE* f = nullptr;
  // arrays already on replica 0's device in the reference layout (awry_index_build): no staging
  const uint64_t* dev_blocks = nullptr;
  const uint64_t* dev_sa = nullptr;
  std::vector<std::pair<const uint8_t*, size_t>> segs;
  size_t cur = 0, off = 0;
  void read(void* dst, size_t n, const char* what) {
    if (n == 0) return;
    if (f) {
      if (n >= (32u << 20)) {
        // large sections (blocks, SA words): positional reads from several threads -- one thread
        // copying out of the page cache tops out at a few GB/s, far below the PCIe rate behind it
        off_t pos = ftello(f);
        int fd = fileno(f);
        unsigned nt = io_threads();
        size_t per = ((n + nt - 1) / nt + 4095) & ~size_t(4095);
        std::vector<std::thread> th;
        std::vector<int> ok(nt, 1);
        for (unsigned t = 0; t < nt; t++) {
          size_t lo = std::min(n, size_t(t) * per), hi = std::min(n, lo + per);
          if (lo >= hi) break;
          th.emplace_back([=, &ok] {
            size_t done = lo;
            while (done < hi) {
              ssize_t r = pread(fd, static_cast<char*>(dst) + done, hi - done, pos + off_t(done));
              if (r <= 0) {
                ok[t] = 0;
                return;
              }
              done += size_t(r);
            }
          });
        }
        for (auto& t : th) t.join();
        for (unsigned t = 0; t < nt; t++)
          if (!ok[t]) fail(AWRY_ERR_IO, "unexpected end of file while reading %s", what);
        if (fseeko(f, pos + off_t(n), SEEK_SET) != 0) fail(AWRY_ERR_IO, "seek failed while reading %s", what);
        return;
      }
      if (fread(dst, 1, n, f) != n) fail(AWRY_ERR_IO, "unexpected end of file while reading %s", what);
      return;
    }
    uint8_t* d = static_cast<uint8_t*>(dst);
    while (n) {
      if (cur >= segs.size()) fail(AWRY_ERR_INVALID_ARG, "index parts too short while reading %s", what);
      size_t take = std::min(segs[cur].second - off, n);
      memcpy(d, segs[cur].first + off, take);
      off += take;
      d += take;
      n -= take;
      if (off == segs[cur].second) {
        cur++;
        off = 0;
      }
    }
  }
};

void set_view_constants(awry_index* ix, Replica& r) {
  IndexView& v = r.view;
  v.blocks = r.d_blocks;
  v.sa_words = r.d_sa;
  v.table = r.d_table;
  v.seq_starts = r.d_seq_starts;
  v.pair_blocks = r.d_pair;
  v.full_sa = r.d_full_sa;
  for (int i = 0; i < 16; i++) v.c2[i] = r.c2[i];
  v.bwt_len = uint32_t(ix->bwt_len);
  v.sa_ratio = uint32_t(ix->sa_ratio);
  v.sa_pow2 = (ix->sa_ratio & (ix->sa_ratio - 1)) == 0 ? 1u : 0u;
  v.sa_ratio_shift = v.sa_pow2 ? uint32_t(__builtin_ctzll(ix->sa_ratio)) : 0u;
  v.sa_bits = ix->sa_bits;
  v.kmer_len = ix->kmer_len_dev;
  v.n_seqs = uint32_t(ix->seq_starts.size());
  v.alphabet = uint32_t(ix->alphabet);
  for (int i = 0; i < 24; i++) v.c_lo[i] = 1, v.c_hi[i] = 0;
  if (ix->alphabet == AWRY_NUCLEOTIDE) {
    static const int ref_of_dsym[6] = {1, 2, 3, 5, 4, 0};  // A C G T N $
    for (int d = 0; d < 6; d++) {
      int ri = ref_of_dsym[d];
      v.c_lo[d] = uint32_t(ix->prefix_sums[ri]);
      v.c_hi[d] = uint32_t(ix->prefix_sums[ri + 1] - 1);
    }
  } else {
    for (int s = 0; s < 22; s++) {
      v.c_lo[s] = uint32_t(ix->prefix_sums[s]);
      v.c_hi[s] = uint32_t(ix->prefix_sums[s + 1] - 1);
    }
  }
}

// Streams the reference-layout blocks through a pinned double buffer and re-lays them out on
// the device; then SA words, then the seed table.
void build_replica0(awry_index* ix, Replica& r, Source& src_blocks_then_rest, bool skip_file_table) {
  DeviceGuard dg(r.device);
  CU(init_device_tables());
  CU(cudaDeviceGetAttribute(&r.sm_count, cudaDevAttrMultiProcessorCount, r.device));
  const uint64_t n_ref_blocks = (ix->bwt_len + 255) / 256;
  const size_t ref_block_bytes = ix->alphabet == AWRY_NUCLEOTIDE ? 160 : 352;
  // nucleotide: 2 x 64-B device blocks per 256-row reference block; amino: 4 x 128-B blocks
  r.bytes_blocks = size_t(n_ref_blocks) * (ix->alphabet == AWRY_NUCLEOTIDE ? 128 : 512);
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_blocks), r.bytes_blocks + 256));
  CU(cudaMemset(r.d_blocks, 0, r.bytes_blocks + 256));
  unsigned int* d_dollar = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&d_dollar), 4));
  CU(cudaMemset(d_dollar, 0xff, 4));

  cudaStream_t st;
  CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const uint64_t CHUNK = 1u << 18;  // reference blocks per staging buffer (40 / 88 MiB)
  uint8_t* h_stage[2] = {nullptr, nullptr};
  uint64_t* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  uint64_t chunk_blocks = std::min<uint64_t>(CHUNK, n_ref_blocks);
  const bool staged = !src_blocks_then_rest.dev_blocks || !src_blocks_then_rest.dev_sa;
  for (int i = 0; i < 2 && staged; i++) {
    CU(cudaHostAlloc(reinterpret_cast<void**>(&h_stage[i]), chunk_blocks * ref_block_bytes, cudaHostAllocDefault));
    CU(cudaMalloc(reinterpret_cast<void**>(&d_stage[i]), chunk_blocks * ref_block_bytes));
    CU(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  int slot = 0;
  if (src_blocks_then_rest.dev_blocks)
    CU(launch_transpose(ix->alphabet, src_blocks_then_rest.dev_blocks, 0, n_ref_blocks, ix->bwt_len, r.d_blocks, d_dollar, st));
  for (uint64_t b0 = 0; b0 < n_ref_blocks && !src_blocks_then_rest.dev_blocks; b0 += chunk_blocks, slot ^= 1) {
    uint64_t nb = std::min(chunk_blocks, n_ref_blocks - b0);
    CU(cudaEventSynchronize(ev[slot]));
    src_blocks_then_rest.read(h_stage[slot], nb * ref_block_bytes, "bwt blocks");
    CU(cudaMemcpyAsync(d_stage[slot], h_stage[slot], nb * ref_block_bytes, cudaMemcpyHostToDevice, st));
    CU(launch_transpose(ix->alphabet, d_stage[slot], b0, nb, ix->bwt_len, r.d_blocks, d_dollar, st));
    CU(cudaEventRecord(ev[slot], st));
  }
  CU(cudaStreamSynchronize(st));
  unsigned int dollar = 0;
  CU(cudaMemcpy(&dollar, d_dollar, 4, cudaMemcpyDeviceToHost));
  cudaFree(d_dollar);
  if (dollar == 0xffffffffu) fail(AWRY_ERR_FORMAT, "no sentinel row found in the BWT blocks");

  // prefix sums (fm_index_file.rs:265-270)
  src_blocks_then_rest.read(ix->prefix_sums, size_t(ix->card + 1) * 8, "prefix sums");
  if (ix->prefix_sums[ix->card] != ix->bwt_len)
    fail(AWRY_ERR_FORMAT, "prefix sums do not add up to bwt_len (%llu vs %llu)",
         (unsigned long long)ix->prefix_sums[ix->card], (unsigned long long)ix->bwt_len);

  // sampled suffix array words, verbatim (fm_index_file.rs:272-278)
  r.bytes_sa = size_t(ix->n_sa_words + 2) * 8;
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_sa), r.bytes_sa));
  CU(cudaMemsetAsync(r.d_sa, 0, r.bytes_sa, st));
  if (src_blocks_then_rest.dev_sa) {
    CU(cudaMemcpyAsync(r.d_sa, src_blocks_then_rest.dev_sa, ix->n_sa_words * 8, cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
  } else {
    const size_t words_per_chunk = chunk_blocks * ref_block_bytes / 8;
    slot = 0;
    for (uint64_t w0 = 0; w0 < ix->n_sa_words; w0 += words_per_chunk, slot ^= 1) {
      uint64_t nw = std::min<uint64_t>(words_per_chunk, ix->n_sa_words - w0);
      CU(cudaEventSynchronize(ev[slot]));
      src_blocks_then_rest.read(h_stage[slot], nw * 8, "sampled suffix array");
      CU(cudaMemcpyAsync(r.d_sa + w0, h_stage[slot], nw * 8, cudaMemcpyHostToDevice, st));
      CU(cudaEventRecord(ev[slot], st));
    }
    CU(cudaStreamSynchronize(st));
  }
  for (int i = 0; i < 2 && staged; i++) {
    cudaFreeHost(h_stage[i]);
    cudaFree(d_stage[i]);
    cudaEventDestroy(ev[i]);
  }

  // k-mer table section of the file: 1 byte k, then (card-2)^k x 16 B that the reference never
  // reads back when searching (kmer_lookup_table.rs:90-110) and that is incomplete.
  if (skip_file_table) {
    uint8_t k = 0;
    src_blocks_then_rest.read(&k, 1, "kmer length");
    ix->kmer_len_file = k;
    uint64_t n_entries = ipow(uint64_t(ix->card - 2), k);
    if (src_blocks_then_rest.f) {
      if (fseeko(src_blocks_then_rest.f, off_t(n_entries * 16), SEEK_CUR) != 0)
        fail(AWRY_ERR_IO, "cannot seek past the kmer table");
    }
  }

  // sequence starts on the device (filled by the caller into ix->seq_starts before this returns
  // for parts; for files they follow the table -- read by the caller after this function)
  r.view.dollar_row = dollar;
  CU(cudaStreamDestroy(st));
}

// The *_device entry points take their scratch from the stream-ordered pool; without a release
// threshold the pool hands the memory back to the driver at every synchronisation and each call
// would re-map ~1 GB.
void keep_pool_memory(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
}

// set by awry_index_build for replicas that exist only to fill a file's k-mer table section
thread_local bool g_skip_accelerators = false;

void finish_replica0(awry_index* ix, Replica& r) {
  DeviceGuard dg(r.device);
  if (ix->seq_starts.empty()) ix->seq_starts.push_back(0);
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_seq_starts), ix->seq_starts.size() * 8));
  CU(cudaMemcpy(r.d_seq_starts, ix->seq_starts.data(), ix->seq_starts.size() * 8, cudaMemcpyHostToDevice));
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_async_flag), 8));
  CU(cudaMemset(r.d_async_flag, 0xff, 8));
  keep_pool_memory(r.device);

  // device seed table: k_dev = min(k_file, cap); the table only accelerates, results are those
  // of the plain backward search either way
  uint32_t cap = ix->alphabet == AWRY_NUCLEOTIDE ? 14 : 6;
  uint32_t k = std::min<uint32_t>(ix->kmer_len_file, cap);
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  while (k > 0 && table_entries(ix->alphabet, k) * 8 > free_b / 4) k--;
  ix->kmer_len_dev = k;
  uint32_t dollar = r.view.dollar_row;
  set_view_constants(ix, r);
  r.view.dollar_row = dollar;
  if (k > 0) {
    r.bytes_table = table_entries(ix->alphabet, k) * 8;
    CU(cudaMalloc(reinterpret_cast<void**>(&r.d_table), r.bytes_table));
    r.view.table = r.d_table;
    IndexView v = r.view;
    v.kmer_len = 0;
    CU(launch_build_table(v, r.d_table, k, nullptr));
    CU(cudaDeviceSynchronize());
  }
  // nucleotide pair index: two query symbols per block access (AWRY_B200_PAIR_INDEX=0 disables)
  const char* env = getenv("AWRY_B200_PAIR_INDEX");
  bool want_pair = ix->alphabet == AWRY_NUCLEOTIDE && !(env && env[0] == '0') && !g_skip_accelerators;
  if (want_pair) {
    size_t bytes = size_t(pair_block_count(ix->bwt_len)) * 128;
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (bytes + bytes / 2 + (1u << 28) < free_b) {
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_pair), bytes + 256));
      CU(build_pair_index(r.view, r.d_pair, r.c2, nullptr));
      r.bytes_pair = bytes;
      r.view.pair_blocks = r.d_pair;
      for (int i = 0; i < 16; i++) r.view.c2[i] = r.c2[i];
    }
  }
  // unsampled suffix array: locate becomes a gather instead of a geometric LF-walk per hit.
  // AWRY_B200_FULL_SA=0 never, =1 whenever it fits, default: when it takes < 1/3 of the free memory.
  const char* fs = getenv("AWRY_B200_FULL_SA");
  if (!(fs && fs[0] == '0') && !g_skip_accelerators && ix->sa_ratio > 1) {
    size_t bytes = size_t(ix->bwt_len) * 4;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const bool force = fs && fs[0] == '1';
    if (force ? bytes + (1u << 28) < free_b : bytes < free_b / 3) {
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_full_sa), bytes + 256));
      CU(build_full_sa(r.view, r.d_full_sa, r.sm_count, nullptr));
      CU(cudaDeviceSynchronize());
      r.bytes_full_sa = bytes;
      r.view.full_sa = r.d_full_sa;
    }
  }
}

void clone_replica(awry_index* ix, const Replica& src, Replica& dst) {
  DeviceGuard dg(dst.device);
  CU(init_device_tables());
  CU(cudaDeviceGetAttribute(&dst.sm_count, cudaDevAttrMultiProcessorCount, dst.device));
  dst.bytes_blocks = src.bytes_blocks;
  dst.bytes_sa = src.bytes_sa;
  dst.bytes_table = src.bytes_table;
  CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_blocks), src.bytes_blocks + 256));
  CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_sa), src.bytes_sa));
  CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_seq_starts), ix->seq_starts.size() * 8));
  CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_async_flag), 8));
  CU(cudaMemset(dst.d_async_flag, 0xff, 8));
  keep_pool_memory(dst.device);
  // fan-out over NVLink instead of N PCIe uploads
  CU(cudaMemcpyPeer(dst.d_blocks, dst.device, src.d_blocks, src.device, src.bytes_blocks + 256));
  CU(cudaMemcpyPeer(dst.d_sa, dst.device, src.d_sa, src.device, src.bytes_sa));
  CU(cudaMemcpyPeer(dst.d_seq_starts, dst.device, src.d_seq_starts, src.device, ix->seq_starts.size() * 8));
  if (src.bytes_table) {
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_table), src.bytes_table));
    CU(cudaMemcpyPeer(dst.d_table, dst.device, src.d_table, src.device, src.bytes_table));
  }
  if (src.bytes_pair) {
    dst.bytes_pair = src.bytes_pair;
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_pair), src.bytes_pair + 256));
    CU(cudaMemcpyPeer(dst.d_pair, dst.device, src.d_pair, src.device, src.bytes_pair));
    memcpy(dst.c2, src.c2, sizeof dst.c2);
  }
  if (src.bytes_full_sa) {
    dst.bytes_full_sa = src.bytes_full_sa;
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_full_sa), src.bytes_full_sa + 256));
    CU(cudaMemcpyPeer(dst.d_full_sa, dst.device, src.d_full_sa, src.device, src.bytes_full_sa));
  }
  uint32_t dollar = src.view.dollar_row;
  set_view_constants(ix, dst);
  dst.view.dollar_row = dollar;
}

std::vector<int> pick_devices(const int* devices, int n_dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(AWRY_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  std::vector<int> out;
  if (devices == nullptr || n_dev <= 0) {
    out.push_back(0);
  } else {
    for (int i = 0; i < n_dev; i++) {
      if (devices[i] < 0 || devices[i] >= count) fail(AWRY_ERR_INVALID_ARG, "device %d out of range (0..%d)", devices[i], count - 1);
      out.push_back(devices[i]);
    }
  }
  for (int d : out) {
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, d));
    if (p.major < 10) fail(AWRY_ERR_CUDA, "device %d is sm_%d%d; this build targets sm_100a (B200)", d, p.major, p.minor);
  }
  return out;
}

void check_header(awry_index* ix) {
  if (ix->alphabet != AWRY_NUCLEOTIDE && ix->alphabet != AWRY_AMINO)
    fail(AWRY_ERR_FORMAT, "invalid symbol alphabet id %d", ix->alphabet);
  if (ix->sa_ratio == 0) fail(AWRY_ERR_FORMAT, "suffix array compression ratio is 0");
  if (ix->bwt_len < 2) fail(AWRY_ERR_FORMAT, "bwt_len %llu too small", (unsigned long long)ix->bwt_len);
  if (ix->bwt_len >= (1ull << 32) - 256)
    fail(AWRY_ERR_UNSUPPORTED, "bwt_len %llu >= 2^32: the device layout uses 32-bit row pointers",
         (unsigned long long)ix->bwt_len);
  if (ix->sa_ratio >= (1ull << 32)) fail(AWRY_ERR_UNSUPPORTED, "suffix array compression ratio too large");
  ix->card = ix->alphabet == AWRY_NUCLEOTIDE ? 6 : 22;
  ix->sa_bits = bits_per_element(ix->bwt_len);
  ix->n_sa_words = sa_word_len(ix->bwt_len, ix->sa_ratio);
}

void make_replicas(awry_index* ix, const std::vector<int>& devs, Source& src, bool from_file) {
  for (int d : devs) {
    auto r = std::make_unique<Replica>();
    r->device = d;
    ix->reps.push_back(std::move(r));
  }
  Replica& r0 = *ix->reps[0];
  build_replica0(ix, r0, src, from_file);
  if (from_file) {
    // sequence index (sequence_index.rs:155-183)
    uint64_t n = 0;
    src.read(&n, 8, "sequence count");
    if (n > (1ull << 32)) fail(AWRY_ERR_FORMAT, "implausible sequence count");
    for (uint64_t i = 0; i < n; i++) {
      uint64_t start = 0, hl = 0;
      src.read(&start, 8, "sequence start");
      src.read(&hl, 8, "header length");
      if (hl > (1ull << 30)) fail(AWRY_ERR_FORMAT, "implausible header length");
      std::string h(hl, '\0');
      src.read(h.data(), hl, "header");
      ix->seq_starts.push_back(start);
      ix->headers.push_back(std::move(h));
    }
  }
  finish_replica0(ix, r0);
  for (size_t i = 1; i < ix->reps.size(); i++) clone_replica(ix, r0, *ix->reps[i]);
}

// ------------------------------------------------------------------ batched pipelines

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

void parallel_memcpy(void* dst, const void* src, size_t n) {
  const size_t MIN_PER_THREAD = 8u << 20;
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned nt = unsigned(std::min<size_t>(std::min(hw, 16u), n / MIN_PER_THREAD));
  if (nt <= 1) {
    memcpy(dst, src, n);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++) {
    size_t lo = n * t / nt, hi = n * (t + 1) / nt;
    th.emplace_back([=] { memcpy(static_cast<char*>(dst) + lo, static_cast<const char*>(src) + lo, hi - lo); });
  }
  for (auto& x : th) x.join();
}

// Pipeline chunk: small enough that the exposed first upload / last kernel are a few percent of a
// 10 M-read batch, large enough (~0.9 M reads) to keep the persistent search grid busy.
constexpr uint64_t CHUNK_MAX_Q = 1u << 20;  // queries per pipeline chunk
uint64_t chunk_max_bytes() {                 // query bytes per pipeline chunk (AWRY_B200_CHUNK_MB, default 128)
  static const uint64_t v = [] {
    uint64_t mb = 128;
    if (const char* e = getenv("AWRY_B200_CHUNK_MB")) mb = std::min<uint64_t>(1024, std::max<uint64_t>(1, strtoull(e, nullptr, 10)));
    return mb << 20;
  }();
  return v;
}

struct Chunk {
  uint64_t q0, q1, b0, b1;
};

std::vector<Chunk> make_chunks(const uint64_t* qoff, uint64_t q_lo, uint64_t q_hi, uint64_t max_q,
                               uint64_t max_bytes) {
  std::vector<Chunk> out;
  uint64_t q = q_lo;
  while (q < q_hi) {
    uint64_t hi = std::min(q_hi, q + max_q);
    // largest hi with qoff[hi] - qoff[q] <= max_bytes (at least one query)
    if (qoff[hi] - qoff[q] > max_bytes) {
      uint64_t lo2 = q + 1, hi2 = hi;
      while (lo2 < hi2) {
        uint64_t mid = (lo2 + hi2 + 1) / 2;
        if (qoff[mid] - qoff[q] <= max_bytes)
          lo2 = mid;
        else
          hi2 = mid - 1;
      }
      hi = lo2;
    }
    out.push_back(Chunk{q, hi, qoff[q], qoff[hi]});
    q = hi;
  }
  return out;
}

// Pipeline fill and drain: the device idles while the host prepares the first chunk, and the host idles
// while the device works on the last one.  Splitting the first chunk into 1/4 + 3/4 and the last into
// 1/2 + 1/4 + 1/4 (by queries) shortens both ends without paying the per-chunk overhead everywhere.
std::vector<Chunk> taper_chunks(std::vector<Chunk> in, const uint64_t* qoff) {
  if (in.size() < 3) return in;
  auto split = [&](const Chunk& c, std::initializer_list<double> cuts, std::vector<Chunk>& dst) {
    uint64_t nq = c.q1 - c.q0, prev = c.q0;
    for (double f : cuts) {
      uint64_t at = c.q0 + uint64_t(double(nq) * f);
      if (at > prev && at < c.q1) {
        dst.push_back(Chunk{prev, at, qoff[prev], qoff[at]});
        prev = at;
      }
    }
    dst.push_back(Chunk{prev, c.q1, qoff[prev], qoff[c.q1]});
  };
  std::vector<Chunk> out;
  const uint64_t MIN_SPLIT = 64u << 20;  // only chunks worth splitting
  if (in.front().b1 - in.front().b0 >= MIN_SPLIT)
    split(in.front(), {0.25}, out);
  else
    out.push_back(in.front());
  for (size_t i = 1; i + 1 < in.size(); i++) out.push_back(in[i]);
  if (in.back().b1 - in.back().b0 >= MIN_SPLIT)
    split(in.back(), {0.5, 0.75}, out);
  else
    out.push_back(in.back());
  return out;
}

bool host_pack_enabled() {
  if (g_host_pack >= 0) return g_host_pack == 1 && host_pack_supported();
  static const bool on = [] {
    if (const char* e = getenv("AWRY_B200_HOST_PACK")) return e[0] != '0' && host_pack_supported();
    return host_pack_supported() && host_pool_threads() >= 4;
  }();
  return on;
}

// Pack on the host or send ASCII?  Packing wins when the host cores pack faster than the PCIe link moves
// bytes (16 threads: 84 GB/s vs 52 GB/s for one GPU) and loses when a process has few cores and shares
// the host's uplinks (4 threads per GPU on an 8-GPU box: measured 690 M reads/s packed vs 1095 M raw).
// Both rates are MEASURED and smoothed across calls: the host clock around the packer (H), CUDA events
// around the copy of a raw first chunk, when nothing else is in flight (P).  H > 1.15 P: pack everything
// (one GPU, 16 threads: mixing raw chunks in was measured and is worse there -- a 128-MiB raw copy holds
// the copy engine for 2.6 ms and starves the search kernel, profiles/r01_s19_e2e_pack_share.log).
// Otherwise host and link are used together: a share f = H / (P + 0.75 H) of the bytes is packed
// (+5 % / +14 % / -8 % at 2 / 4 / 8 GPUs of a 32-core box, profiles/r01_s23_e2e_mixed_multi_gpu.log).
// AWRY_B200_PACK_SHARE=<0..1> pins the packed share of the bytes instead (experiments).
struct PackBalance {
  std::mutex mu;
  double host_rate = 0, link_rate = 0;  // bytes/s, 0 = not measured yet
  double fixed_share = -1;
  uint64_t calls = 0;
  bool mixed = true;  // AWRY_B200_PACK_MIXED=0: all-or-nothing
  PackBalance() {
    if (const char* e = getenv("AWRY_B200_PACK_SHARE")) fixed_share = std::min(1.0, std::max(0.0, atof(e)));
    if (const char* e = getenv("AWRY_B200_PACK_MIXED")) mixed = e[0] != '0';
  }
  void note_host(double bytes, double seconds) {
    if (seconds <= 0 || bytes < (8 << 20)) return;
    std::lock_guard<std::mutex> lk(mu);
    double r = bytes / seconds;
    host_rate = host_rate > 0 ? 0.7 * host_rate + 0.3 * r : r;
  }
  void note_link(double bytes, double seconds) {
    if (seconds <= 0 || bytes < (8 << 20)) return;
    std::lock_guard<std::mutex> lk(mu);
    double r = bytes / seconds;
    link_rate = link_rate > 0 ? 0.7 * link_rate + 0.3 * r : r;
  }
  // per call: the packed share of the bytes and whether chunk 0 / chunk 1 serve as probes
  struct Plan {
    double share;
    bool probe_link, probe_host;
  };
  Plan plan() {
    std::lock_guard<std::mutex> lk(mu);
    if (fixed_share >= 0) return Plan{fixed_share, false, false};
    const bool refresh = calls++ % 32 == 0;
    if (host_rate <= 0 || link_rate <= 0) return Plan{1.0, true, true};
    if (host_rate > 1.15 * link_rate) return Plan{1.0, refresh, false};
    // the host is not clearly faster than the link: use both (see the share formula above)
    double f = mixed ? host_rate / (link_rate + 0.75 * host_rate) : 0.0;
    if (f < 0.15) f = 0.0;
    return Plan{f, true, f == 0.0 && refresh};
  }
};
PackBalance g_balance;

void validate_offsets(const uint64_t* qoff, uint64_t nq) {
  // cheap sanity check on the ends; per-query monotonicity is checked on the device (prepass)
  if (nq && qoff[nq] < qoff[0]) fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone");
}

// Uploads one chunk of queries and runs prepass + search on ws->st.  The search result lands
// in ws->d_out in the requested mode.
// `may_pack`: the caller's say on host packing for this chunk (see raw_chunk_period)
void enqueue_search(const awry_index* ix, Replica& r, Workspace* ws, const uint8_t* qbytes,
                    const uint64_t* qoff, const Chunk& c, SearchOut mode, bool src_pinned, bool may_pack = true,
                    bool probe_link = false) {
  const uint64_t nq = c.q1 - c.q0, nbytes = c.b1 - c.b0;
  const size_t out_elem = mode == OUT_COUNT_U64 ? 8 : mode == OUT_RANGE_U64 ? 16 : 8;
  Workspace::grow_dev(ws->d_qbytes, ws->d_qbytes_cap, size_t(nbytes) + 16);
  Workspace::grow_dev(ws->d_qoff, ws->d_qoff_cap, size_t(nq) + 1);
  Workspace::grow_dev(ws->d_qwords, ws->d_qwords_cap, size_t(packed_words(ix->alphabet, nq, nbytes)));
  Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * out_elem);
  Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, size_t(nq) + 2);
  const uint8_t* src_b = qbytes + c.b0;
  const uint64_t* src_o = qoff + c.q0;
  // Nucleotide chunks are packed to 2 bits per base by the host cores before the copy (a quarter of
  // the PCIe bytes; works the same for pageable and pinned caller memory).  Chunks with many bytes
  // outside ACGT go up as ASCII.
  bool packed = false;
  ws->link_probe_bytes = 0;
  if (may_pack && ix->alphabet == AWRY_NUCLEOTIDE && host_pack_enabled() && nbytes >= 4096) {
    Workspace::grow_host(ws->h_qbytes, ws->h_qbytes_cap, size_t(nbytes) / 4 + 64);
    auto t0 = std::chrono::steady_clock::now();
    packed = host_pack_dna(src_b, size_t(nbytes), ws->h_qbytes, ws->exc_tmp, 64);
    if (packed) g_balance.note_host(double(nbytes), std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  if (packed) {
    const size_t n_exc = ws->exc_tmp.size();
    const size_t pbytes = (size_t(nbytes) + 3) / 4;
    if (!src_pinned) {  // offsets: staged through pinned memory by the pool
      Workspace::grow_host(ws->h_qoff, ws->h_qoff_cap, size_t(nq) + 1);
      parallel_memcpy(ws->h_qoff, src_o, (nq + 1) * 8);
      src_o = ws->h_qoff;
    }
    CU(cudaMemcpyAsync(ws->d_qbytes, ws->h_qbytes, pbytes + 8, cudaMemcpyHostToDevice, ws->st));
    CU(cudaMemcpyAsync(ws->d_qoff, src_o, (nq + 1) * 8, cudaMemcpyHostToDevice, ws->st));
    if (n_exc) {
      Workspace::grow_host(ws->h_exc, ws->h_exc_cap, n_exc);
      Workspace::grow_dev(ws->d_exc, ws->d_exc_cap, n_exc);
      memcpy(ws->h_exc, ws->exc_tmp.data(), n_exc * 8);
      CU(cudaMemcpyAsync(ws->d_exc, ws->h_exc, n_exc * 8, cudaMemcpyHostToDevice, ws->st));
    }
    g_prof.h2d += pbytes + 8 + (nq + 1) * 8 + n_exc * 8;
    CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, ws->st));
    {
      ProfScope p(2, r.device, ws->st);
      CU(launch_pack2(reinterpret_cast<const uint32_t*>(ws->d_qbytes), c.b0, ws->d_qoff, nq,
                      ws->d_qwords - 4 * (c.b0 >> packed_unit_shift(ix->alphabet)), ws->d_exc, n_exc, ws->d_flag, ws->st));
    }
  } else {
    if (!src_pinned) {
      Workspace::grow_host(ws->h_qbytes, ws->h_qbytes_cap, size_t(nbytes) + 16);
      Workspace::grow_host(ws->h_qoff, ws->h_qoff_cap, size_t(nq) + 1);
      parallel_memcpy(ws->h_qbytes, src_b, nbytes);
      parallel_memcpy(ws->h_qoff, src_o, (nq + 1) * 8);
      src_b = ws->h_qbytes;
      src_o = ws->h_qoff;
    }
    const bool probe = probe_link && src_pinned && nbytes >= (8u << 20);
    if (probe) {
      if (!ws->ev_a) {
        CU(cudaEventCreate(&ws->ev_a));
        CU(cudaEventCreate(&ws->ev_b));
      }
      CU(cudaEventRecord(ws->ev_a, ws->st));
    }
    if (nbytes) CU(cudaMemcpyAsync(ws->d_qbytes, src_b, nbytes, cudaMemcpyHostToDevice, ws->st));
    if (probe) {
      CU(cudaEventRecord(ws->ev_b, ws->st));
      ws->link_probe_bytes = nbytes;
    }
    CU(cudaMemcpyAsync(ws->d_qoff, src_o, (nq + 1) * 8, cudaMemcpyHostToDevice, ws->st));
    g_prof.h2d += nbytes + (nq + 1) * 8;
    CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, ws->st));
    // offsets stay absolute: the kernels subtract the chunk's byte base
    {
      ProfScope p(2, r.device, ws->st);
      CU(launch_pack(ix->alphabet, ws->d_qbytes - c.b0, ws->d_qoff, nq, ws->d_qwords - 4 * (c.b0 >> packed_unit_shift(ix->alphabet)), ws->d_flag, ws->st));
    }
  }
  {
    ProfScope p(0, r.device, ws->st);
    SearchVariant v = g_variant;
    v.avg_len = uint32_t(std::min<uint64_t>(nbytes / std::max<uint64_t>(1, nq), 1u << 30));
    CU(launch_search(r.view, ws->d_qwords - 4 * (c.b0 >> packed_unit_shift(ix->alphabet)), ws->d_qoff, nq, mode, ws->d_out, ws->d_defer, v, r.sm_count, ws->st));
  }
  CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, ws->st));
}

void check_flag(Workspace* ws, const Chunk& c) {
  if (*ws->h_flag != ~0ull)
    fail(AWRY_ERR_INVALID_QUERY,
         "query %llu is empty or contains a sentinel ('$'/'#'): the reference panics on it "
         "(fm_index.rs:406, bwt.rs:127)",
         (unsigned long long)(c.q0 + *ws->h_flag));
}

// count / range search over [q_lo, q_hi) on one replica, 3-deep pipeline
void search_on_replica(const awry_index* ix, Replica& r, const uint8_t* qbytes, const uint64_t* qoff,
                       uint64_t q_lo, uint64_t q_hi, SearchOut mode, void* out) {
  if (q_lo >= q_hi) return;
  DeviceGuard dg(r.device);
  const size_t out_elem = mode == OUT_COUNT_U64 ? 8 : 16;
  const bool src_pinned = is_pinned(qbytes) && is_pinned(qoff);
  const bool dst_pinned = is_pinned(out);
  auto chunks = taper_chunks(make_chunks(qoff, q_lo, q_hi, CHUNK_MAX_Q, chunk_max_bytes()), qoff);
  constexpr int DEPTH = 3;
  Workspace* ws[DEPTH] = {nullptr, nullptr, nullptr};
  int pending[DEPTH] = {-1, -1, -1};
  auto finish = [&](int s) {
    if (pending[s] < 0) return;
    const Chunk& c = chunks[size_t(pending[s])];
    CU(cudaEventSynchronize(ws[s]->done));
    if (ws[s]->link_probe_bytes) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ws[s]->ev_a, ws[s]->ev_b) == cudaSuccess)
        g_balance.note_link(double(ws[s]->link_probe_bytes), double(ms) * 1e-3);
      ws[s]->link_probe_bytes = 0;
    }
    check_flag(ws[s], c);
    if (!dst_pinned)
      parallel_memcpy(static_cast<char*>(out) + c.q0 * out_elem, ws[s]->h_out, (c.q1 - c.q0) * out_elem);
    pending[s] = -1;
  };
  uint64_t bytes_total = 0, bytes_packed = 0;  // of the chunks enqueued so far (pinned sources only)
  const bool balanced = src_pinned && ix->alphabet == AWRY_NUCLEOTIDE && host_pack_enabled();
  PackBalance::Plan plan{1.0, false, false};
  if (balanced) plan = g_balance.plan();
  try {
    for (size_t i = 0; i < chunks.size(); i++) {
      int s = int(i % DEPTH);
      if (!ws[s]) ws[s] = r.acquire();
      finish(s);
      const Chunk& c = chunks[i];
      bool may_pack = true, probe = false;
      if (balanced) {
        if (i == 0 && plan.probe_link && chunks.size() > 1) {
          may_pack = false;  // nothing else is in flight: the cleanest moment to time the link
          probe = true;
        } else if (i == 1 && plan.probe_host) {
          may_pack = true;
        } else {
          may_pack = double(bytes_packed) < plan.share * double(bytes_total + (c.b1 - c.b0));
        }
        bytes_total += c.b1 - c.b0;
        if (may_pack) bytes_packed += c.b1 - c.b0;
      }
      enqueue_search(ix, r, ws[s], qbytes, qoff, c, mode, src_pinned, may_pack, probe);
      size_t bytes = (c.q1 - c.q0) * out_elem;
      void* dst = static_cast<char*>(out) + c.q0 * out_elem;
      if (!dst_pinned) {
        Workspace::grow_host(ws[s]->h_out, ws[s]->h_out_cap, bytes);
        dst = ws[s]->h_out;
      }
      CU(cudaMemcpyAsync(dst, ws[s]->d_out, bytes, cudaMemcpyDeviceToHost, ws[s]->st));
      g_prof.d2h += bytes;
      CU(cudaEventRecord(ws[s]->done, ws[s]->st));
      pending[s] = int(i);
    }
    for (int s = 0; s < DEPTH; s++) finish(s);
  } catch (...) {
    for (int s = 0; s < DEPTH; s++)
      if (ws[s]) {
        cudaStreamSynchronize(ws[s]->st);
        r.release(ws[s]);
      }
    throw;
  }
  for (int s = 0; s < DEPTH; s++)
    if (ws[s]) r.release(ws[s]);
}

// splits [0,nq) across replicas by query bytes; one host thread per replica
template <class F>
void for_each_replica_range(const awry_index* ix, const uint64_t* qoff, uint64_t nq, F&& fn) {
  size_t nr = ix->reps.size();
  if (nr == 1 || nq < 2 * nr) {
    fn(0, 0, nq);
    return;
  }
  std::vector<uint64_t> cut(nr + 1, 0);
  cut[nr] = nq;
  uint64_t total = qoff[nq] - qoff[0];
  for (size_t i = 1; i < nr; i++) {
    uint64_t target = qoff[0] + total * i / nr;
    cut[i] = uint64_t(std::lower_bound(qoff, qoff + nq, target) - qoff);
    cut[i] = std::max(cut[i], cut[i - 1]);
  }
  std::vector<std::thread> th;
  std::vector<int> codes(nr, 0);
  std::vector<std::string> msgs(nr);
  for (size_t i = 0; i < nr; i++)
    th.emplace_back([&, i] {
      try {
        fn(i, cut[i], cut[i + 1]);
      } catch (const ApiError& e) {
        codes[i] = e.code;
        msgs[i] = e.what();
      } catch (const std::exception& e) {
        codes[i] = AWRY_ERR_INVALID_ARG;
        msgs[i] = e.what();
      }
    });
  for (auto& t : th) t.join();
  for (size_t i = 0; i < nr; i++)
    if (codes[i]) fail(codes[i], "%s", msgs[i].c_str());
}

struct LocatePart {
  std::vector<uint64_t> hit_off;  // local CSR over the replica's queries, size n+1
  awry_hit* hits = nullptr;       // malloc'd, or the caller's buffer when ext_cap != 0
  uint64_t n_hits = 0;
  uint64_t ext_cap = 0;           // caller-owned output: capacity in hits (0 = library allocates)
  uint64_t* ext_off = nullptr;    // caller-owned CSR offsets to fill directly (single replica)
};

// CSR offsets of a searched chunk (ws->d_out holds (sp, count) per query), no synchronisation
void locate_chunk_scan(Workspace* ws, uint64_t nq, uint64_t* d_hit_off, cudaStream_t st) {
  const uint2* d_sp_cnt = reinterpret_cast<const uint2*>(ws->d_out);
  size_t temp = 0;
  CU(scan_hit_offsets(d_sp_cnt, nq, d_hit_off, nullptr, temp, st));
  Workspace::grow_dev(reinterpret_cast<uint8_t*&>(ws->d_temp), ws->d_temp_cap, temp + 16);
  CU(scan_hit_offsets(d_sp_cnt, nq, d_hit_off, ws->d_temp, temp, st));
}

// device-side two-pass locate of a chunk whose queries were already searched (ws->d_out holds
// (sp, count) per query).  Step 1: CSR offsets + hit total (one synchronisation).
uint64_t locate_chunk_count(Replica& r, Workspace* ws, uint64_t nq, uint64_t* d_hit_off, cudaStream_t st) {
  (void)r;
  const uint2* d_sp_cnt = reinterpret_cast<const uint2*>(ws->d_out);
  size_t temp = 0;
  CU(scan_hit_offsets(d_sp_cnt, nq, d_hit_off, nullptr, temp, st));
  Workspace::grow_dev(reinterpret_cast<uint8_t*&>(ws->d_temp), ws->d_temp_cap, temp + 16);
  CU(scan_hit_offsets(d_sp_cnt, nq, d_hit_off, ws->d_temp, temp, st));
  uint64_t n_hits = 0;
  CU(cudaMemcpyAsync(&n_hits, d_hit_off + nq, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return n_hits;
}

// Step 2: LF-walk every hit; returns a stream-ordered device buffer with n_hits awry_hit entries.
uint64_t* locate_chunk_walk(Replica& r, Workspace* ws, uint64_t nq, uint64_t n_hits, uint32_t flags,
                            const uint64_t* d_hit_off, cudaStream_t st) {
  if (n_hits == 0) return nullptr;
  IndexView view = r.view;
  if (g_locate_variant == 1) view.full_sa = nullptr;
  const uint2* d_sp_cnt = reinterpret_cast<const uint2*>(ws->d_out);
  uint64_t* d_hits = nullptr;
  CU(cudaMallocAsync(reinterpret_cast<void**>(&d_hits), n_hits * 16 + 16, st));  // pool: no driver round trip
  try {
    if (flags & AWRY_LOCATE_SORTED) {
      uint64_t *d_locs = nullptr, *d_sorted = nullptr;
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d_locs), n_hits * 8 + 16, st));
      CU(cudaMallocAsync(reinterpret_cast<void**>(&d_sorted), n_hits * 8, st));
      {
        ProfScope p(1, r.device, st);
        CU(launch_walk(view, d_sp_cnt, d_hit_off, nq, n_hits, nullptr, d_locs, r.sm_count, st));
      }
      size_t t2 = 0;
      CU(sort_hit_segments(d_locs, d_sorted, n_hits, nq, d_hit_off, nullptr, t2, st));
      Workspace::grow_dev(reinterpret_cast<uint8_t*&>(ws->d_temp), ws->d_temp_cap, t2 + 16);
      CU(sort_hit_segments(d_locs, d_sorted, n_hits, nq, d_hit_off, ws->d_temp, t2, st));
      CU(launch_map_locations(view, d_sorted, n_hits, d_hits, st));
      cudaFreeAsync(d_locs, st);
      cudaFreeAsync(d_sorted, st);
    } else {
      ProfScope p(1, r.device, st);
      CU(launch_walk(view, d_sp_cnt, d_hit_off, nq, n_hits, d_hits, nullptr, r.sm_count, st));
    }
  } catch (...) {
    cudaFreeAsync(d_hits, st);
    throw;
  }
  return d_hits;
}

uint64_t* locate_chunk_device(const awry_index* ix, Replica& r, Workspace* ws, uint64_t nq,
                              uint32_t flags, uint64_t* d_hit_off, uint64_t* n_hits_out,
                              cudaStream_t st) {
  (void)ix;
  *n_hits_out = locate_chunk_count(r, ws, nq, d_hit_off, st);
  return locate_chunk_walk(r, ws, nq, *n_hits_out, flags, d_hit_off, st);
}

// parallel_locate over [q_lo, q_hi) on one replica.  The two passes of a chunk are separated by one small
// device->host read (the hit total sizes pass 2), so chunks are kept small (256 k queries) and three are
// in flight: while the host waits for chunk i's total, chunk i+1 is being packed, copied and searched,
// and chunk i-1's hits are on their way back.
constexpr uint64_t LOCATE_CHUNK_Q = 1u << 18;
constexpr uint64_t LOCATE_CHUNK_BYTES = 32u << 20;

void locate_on_replica(const awry_index* ix, Replica& r, const uint8_t* qbytes, const uint64_t* qoff,
                       uint64_t q_lo, uint64_t q_hi, uint32_t flags, LocatePart& part) {
  const bool ext = part.ext_cap != 0 || part.ext_off != nullptr;
  if (!part.ext_off) part.hit_off.assign(q_hi - q_lo + 1, 0);
  uint64_t* off_base = part.ext_off ? part.ext_off : part.hit_off.data();
  if (q_lo >= q_hi) return;
  DeviceGuard dg(r.device);
  const bool src_pinned = is_pinned(qbytes) && is_pinned(qoff);
  auto chunks = make_chunks(qoff, q_lo, q_hi, LOCATE_CHUNK_Q, std::min(chunk_max_bytes(), LOCATE_CHUNK_BYTES));
  constexpr int DEPTH = 3;
  struct Slot {
    Workspace* ws = nullptr;
    int chunk = -1;
    int phase = 0;  // 1 = searched + scanned (total on its way), 2 = pass 2 enqueued (results on their way)
    uint64_t base = 0;
  } slot[DEPTH];
  size_t cap = 0;
  static const bool trace = getenv("AWRY_B200_TRACE") != nullptr;  // host-side stage times on stderr
  const auto t_origin = std::chrono::steady_clock::now();
  auto mark = [&](const char* what, int i) {
    if (trace)
      fprintf(stderr, "[locate] %8.3f ms  %s %d\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_origin).count(), what, i);
  };
  // offsets of a chunk are rebased onto the replica-local hit count before it; the slot shared with the
  // next chunk (index nq) is written by that chunk, the very last one after the loop
  auto stage_c = [&](Slot& s) {
    if (s.phase != 2) return;
    const Chunk& c = chunks[size_t(s.chunk)];
    CU(cudaEventSynchronize(s.ws->done));
    uint64_t* dst_off = off_base + (c.q0 - q_lo);
    const uint64_t nq = c.q1 - c.q0, base = s.base;
    if (base)
      for (uint64_t i = 0; i < nq; i++) dst_off[i] += base;
    s.phase = 0;
    s.chunk = -1;
  };
  auto stage_a = [&](int i) {
    Slot& s = slot[i % DEPTH];
    if (!s.ws) s.ws = r.acquire();
    stage_c(s);
    Workspace* ws = s.ws;
    const Chunk& c = chunks[size_t(i)];
    const uint64_t nq = c.q1 - c.q0;
    mark("A begin", i);
    enqueue_search(ix, r, ws, qbytes, qoff, c, OUT_SP_CNT_U32, src_pinned);
    mark("A searched", i);
    Workspace::grow_dev(ws->d_hit_off, ws->d_hit_off_cap, size_t(nq) + 1);
    locate_chunk_scan(ws, nq, ws->d_hit_off, ws->st);
    CU(cudaMemcpyAsync(ws->h_total, ws->d_hit_off + nq, 8, cudaMemcpyDeviceToHost, ws->st));
    CU(cudaEventRecord(ws->done, ws->st));
    mark("A end", i);
    s.chunk = i;
    s.phase = 1;
  };
  auto stage_b = [&](int i) {
    Slot& s = slot[i % DEPTH];
    Workspace* ws = s.ws;
    const Chunk& c = chunks[size_t(i)];
    const uint64_t nq = c.q1 - c.q0;
    mark("B wait", i);
    CU(cudaEventSynchronize(ws->done));
    mark("B total known", i);
    const uint64_t n_hits = *ws->h_total;
    check_flag(ws, c);
    s.base = part.n_hits;
    CU(cudaMemcpyAsync(off_base + (c.q0 - q_lo), ws->d_hit_off, nq * 8, cudaMemcpyDeviceToHost, ws->st));
    g_prof.d2h += nq * 8;
    const bool fits = !ext || part.n_hits + n_hits <= part.ext_cap;
    if (n_hits && fits) {
      if (!ext && part.n_hits + n_hits > cap) {
        for (auto& o : slot) stage_c(o);  // copies into the old buffer must land before it moves
        cap = std::max<size_t>(size_t(part.n_hits + n_hits), cap * 2);
        void* np = realloc(part.hits, cap * sizeof(awry_hit));
        if (!np) fail(AWRY_ERR_NOMEM, "out of host memory for %llu hits", (unsigned long long)cap);
        part.hits = static_cast<awry_hit*>(np);
      }
      uint64_t* d_hits = locate_chunk_walk(r, ws, nq, n_hits, flags, ws->d_hit_off, ws->st);
      CU(cudaMemcpyAsync(part.hits + part.n_hits, d_hits, n_hits * 16, cudaMemcpyDeviceToHost, ws->st));
      cudaFreeAsync(d_hits, ws->st);
      g_prof.d2h += n_hits * 16;
    }
    CU(cudaEventRecord(ws->done, ws->st));
    mark("B end", i);
    s.phase = 2;
    part.n_hits += n_hits;  // keeps counting past the capacity so the caller learns the need
  };
  try {
    stage_a(0);
    for (int i = 0; i < int(chunks.size()); i++) {
      if (i + 1 < int(chunks.size())) stage_a(i + 1);
      stage_b(i);
    }
    for (auto& s : slot) stage_c(s);
    off_base[q_hi - q_lo] = part.n_hits;
    mark("done", int(chunks.size()));
  } catch (...) {
    for (auto& s : slot)
      if (s.ws) {
        cudaStreamSynchronize(s.ws->st);
        r.release(s.ws);
      }
    if (!ext) free(part.hits);
    part.hits = nullptr;
    throw;
  }
  for (auto& s : slot)
    if (s.ws) r.release(s.ws);
}

// ------------------------------------------------------------------ streaming reads-file front-end
// FASTQ / FASTA file of queries -> parallel_count / parallel_locate without a host-side parser:
// a reader thread preads the file into a ring of pinned buffers, the raw bytes are uploaded and
// parsed on the device (reads.cu), and the parsed (query bytes, CSR offsets) feed the same pack /
// search / locate kernels as the *_device entry points.  A record cut by a chunk boundary is carried
// into the next chunk by the host.

struct ReadsOut {
  std::vector<uint64_t> counts;
  std::vector<uint64_t> hit_off;  // locate: CSR, n_reads + 1
  awry_hit* hits = nullptr;       // locate: malloc'd
  uint64_t n_hits = 0, hits_cap = 0;
  uint64_t n_reads = 0, n_bases = 0, file_bytes = 0;
};

void parallel_pread(int fd, void* dst, size_t n, off_t pos, const char* what) {
  unsigned nt = n >= (4u << 20) ? io_threads() : 1u;
  size_t per = ((n + nt - 1) / nt + 4095) & ~size_t(4095);
  std::vector<int> ok(nt, 1);
  auto work = [&](unsigned t) {
    size_t lo = std::min(n, size_t(t) * per), hi = std::min(n, lo + per);
    while (lo < hi) {
      ssize_t r = pread(fd, static_cast<char*>(dst) + lo, hi - lo, pos + off_t(lo));
      if (r <= 0) {
        ok[t] = 0;
        return;
      }
      lo += size_t(r);
    }
  };
  if (nt == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  for (unsigned t = 0; t < nt; t++)
    if (!ok[t]) fail(AWRY_ERR_IO, "read error in %s", what);
}

void run_reads_file(const awry_index* ix, const char* path, bool locate, uint32_t flags, ReadsOut& out) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) fail(AWRY_ERR_IO, "cannot open %s: %s", path, strerror(errno));
  struct FdCloser {
    int fd;
    ~FdCloser() { close(fd); }
  } fdc{fd};
  struct stat sb;
  if (fstat(fd, &sb) != 0) fail(AWRY_ERR_IO, "cannot stat %s", path);
  const uint64_t fsize = uint64_t(sb.st_size);
  out.file_bytes = fsize;
  // format: first non-blank byte (of the inflated stream when the file is gzip-compressed)
  int fastq = -1;
  uint64_t data_start = 0;
  gzFile gz = nullptr;
  struct GzCloser {
    gzFile& g;
    ~GzCloser() {
      if (g) gzclose(g);
    }
  } gzc{gz};
  {
    unsigned char head[4096];
    ssize_t got = pread(fd, head, sizeof head, 0);
    bool at_end = fsize <= uint64_t(std::max<ssize_t>(got, 0));
    if (got >= 2 && head[0] == 0x1f && head[1] == 0x8b) {
      gz = gzopen(path, "rb");
      if (!gz) fail(AWRY_ERR_IO, "cannot open %s as a gzip stream", path);
      gzbuffer(gz, 4u << 20);
      got = gzread(gz, head, sizeof head);
      int zerr = Z_OK;
      gzerror(gz, &zerr);  // a truncated stream returns the bytes it has and only flags the error
      if (got < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) fail(AWRY_ERR_FORMAT, "%s: corrupt gzip stream", path);
      at_end = got < ssize_t(sizeof head);
    }
    for (ssize_t i = 0; i < got; i++) {
      unsigned char c = head[i];
      if (c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
      fastq = c == '@' ? 1 : c == '>' ? 0 : -1;
      data_start = uint64_t(i);
      break;
    }
    if (fastq < 0) {
      if (got <= 0 || data_start == 0) {
        bool blank = true;
        for (ssize_t i = 0; i < got; i++) blank &= (head[i] == '\n' || head[i] == '\r' || head[i] == ' ' || head[i] == '\t');
        if (blank && at_end) {  // empty file: zero reads
          if (locate) out.hit_off.assign(1, 0);
          return;
        }
      }
      fail(AWRY_ERR_FORMAT, "%s is neither FASTQ ('@') nor FASTA ('>')", path);
    }
    if (gz && gzseek(gz, z_off_t(data_start), SEEK_SET) < 0) fail(AWRY_ERR_FORMAT, "%s: corrupt gzip stream", path);
  }
  uint64_t CHUNK = 64ull << 20;
  if (const char* e = getenv("AWRY_B200_READS_CHUNK")) CHUNK = std::max<uint64_t>(64, strtoull(e, nullptr, 10));
  CHUNK = std::min<uint64_t>(CHUNK, 512ull << 20);
  // the largest record that can straddle a chunk boundary (AWRY_B200_READS_CARRY, default min(chunk, 16 MiB))
  uint64_t CARRY = std::min<uint64_t>(CHUNK, 16ull << 20);
  if (const char* e = getenv("AWRY_B200_READS_CARRY")) CARRY = std::min<uint64_t>(std::max<uint64_t>(64, strtoull(e, nullptr, 10)), 1ull << 30);
  constexpr int NBUF = 3;

  Replica& r = *ix->reps[0];
  DeviceGuard dg(r.device);
  Workspace* ws = r.acquire();
  cudaStream_t st = ws->st;
  Workspace::ReadsScratch& rs = ws->rs;

  // reader thread state
  std::mutex mu;
  std::condition_variable cv;
  struct Slot {
    bool filled = false;
    uint64_t n = 0;
    bool eof = false;
  } slots[NBUF];
  bool stop = false;
  std::string reader_err;
  std::thread reader;

  auto cleanup = [&] {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    if (reader.joinable()) reader.join();
    cudaStreamSynchronize(st);
    r.release(ws);
  };
  try {
    const uint32_t max_bytes = uint32_t(CARRY + CHUNK + 1);
    if (rs.chunk != CHUNK || rs.carry != CARRY) {
      rs.release();
      for (auto& b : rs.h_buf) CU(cudaHostAlloc(reinterpret_cast<void**>(&b), CARRY + CHUNK + 64, cudaHostAllocDefault));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_raw), max_bytes + 64));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_qbytes), max_bytes + 64));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_nl), size_t(max_bytes) * 4 + 64));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_small), 64));
      rs.temp_bytes = reads_temp_bytes(max_bytes);
      CU(cudaMalloc(&rs.d_temp, rs.temp_bytes));
      CU(cudaHostAlloc(&rs.h_plan, sizeof(ReadsPlan) + 64, cudaHostAllocDefault));
      rs.chunk = CHUNK;
      rs.carry = CARRY;
    }
    uint8_t** h_buf = rs.h_buf;
    uint8_t *d_raw = rs.d_raw, *d_qbytes = rs.d_qbytes;
    uint32_t *d_nl = rs.d_nl, *d_small = rs.d_small;
    uint32_t *&d_seq_len = rs.d_seq_len, *&d_is_hdr = rs.d_is_hdr, *&d_hdr_rank = rs.d_hdr_rank;
    uint64_t *&d_seq_off = rs.d_seq_off, *&d_qoff = rs.d_qoff;
    size_t &cap_lines = rs.cap_lines, &cap_qoff = rs.cap_qoff;
    void* d_temp = rs.d_temp;
    const size_t temp_bytes = rs.temp_bytes;
    ReadsPlan* h_plan = static_cast<ReadsPlan*>(rs.h_plan);
    uint32_t* h_n_lines = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(h_plan) + sizeof(ReadsPlan));

    reader = std::thread([&] {
      uint64_t pos = data_start;
      for (int c = 0;; c++) {
        Slot& s = slots[c % NBUF];
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || !s.filled; });
          if (stop) return;
        }
        uint64_t n = gz ? 0 : std::min<uint64_t>(CHUNK, fsize - pos);
        bool gz_end = false;
        try {
          if (gz) {  // inflate on this thread (one zlib stream: ~0.3-0.5 GB/s); the device side is unchanged
            while (n < CHUNK) {
              int r = gzread(gz, h_buf[c % NBUF] + CARRY + n, unsigned(std::min<uint64_t>(CHUNK - n, 1u << 30)));
              int zerr = Z_OK;
              if (r <= 0) gzerror(gz, &zerr);
              if (r < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) fail(AWRY_ERR_FORMAT, "%s: corrupt or truncated gzip stream", path);
              if (r == 0) {
                gz_end = true;
                break;
              }
              n += uint64_t(r);
            }
          } else if (n) {
            parallel_pread(fd, h_buf[c % NBUF] + CARRY, size_t(n), off_t(pos), path);
          }
        } catch (const ApiError& e) {
          std::lock_guard<std::mutex> lk(mu);
          reader_err = e.what();
          s.filled = true;
          s.n = 0;
          s.eof = true;
          cv.notify_all();
          return;
        }
        pos += n;
        const bool at_eof = gz ? gz_end : pos >= fsize;
        {
          std::lock_guard<std::mutex> lk(mu);
          s.n = n;
          s.eof = at_eof;
          s.filled = true;
        }
        cv.notify_all();
        if (at_eof) return;
      }
    });

    if (locate) out.hit_off.assign(1, 0);
    const int sh = packed_unit_shift(ix->alphabet);
    uint64_t tail_len = 0;
    const uint8_t* tail_src = nullptr;
    int prev_slot = -1;
    for (int c = 0;; c++) {
      const int si = c % NBUF;
      Slot& s = slots[si];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return s.filled; });
      }
      if (!reader_err.empty()) fail(AWRY_ERR_IO, "%s", reader_err.c_str());
      uint8_t* base = h_buf[si] + CARRY - tail_len;
      if (tail_len) memcpy(base, tail_src, tail_len);
      if (prev_slot >= 0) {  // the previous buffer is free once its tail has moved
        {
          std::lock_guard<std::mutex> lk(mu);
          slots[prev_slot].filled = false;
        }
        cv.notify_all();
      }
      uint64_t n = tail_len + s.n;
      const bool eof = s.eof;
      if (eof && n && base[n - 1] != '\n') base[n++] = '\n';
      if (n == 0) break;
      CU(cudaMemcpyAsync(d_raw, base, n, cudaMemcpyHostToDevice, st));
      g_prof.h2d += n;
      CU(reads_find_lines(d_raw, uint32_t(n), d_nl, d_small, d_temp, temp_bytes, st));
      CU(cudaMemcpyAsync(h_n_lines, d_small, 4, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      const uint32_t n_lines = *h_n_lines;
      if (size_t(n_lines) + 2 > cap_lines) {
        cudaFree(d_seq_len);
        cudaFree(d_is_hdr);
        cudaFree(d_hdr_rank);
        cudaFree(d_seq_off);
        d_seq_len = d_is_hdr = d_hdr_rank = nullptr;
        d_seq_off = nullptr;
        cap_lines = size_t(n_lines) + size_t(n_lines) / 4 + 1024;
        CU(cudaMalloc(reinterpret_cast<void**>(&d_seq_len), cap_lines * 4));
        CU(cudaMalloc(reinterpret_cast<void**>(&d_is_hdr), cap_lines * 4));
        CU(cudaMalloc(reinterpret_cast<void**>(&d_hdr_rank), cap_lines * 4));
        CU(cudaMalloc(reinterpret_cast<void**>(&d_seq_off), cap_lines * 8));
      }
      if (size_t(n_lines) + 2 > cap_qoff) {
        cudaFree(d_qoff);
        d_qoff = nullptr;
        cap_qoff = size_t(n_lines) + size_t(n_lines) / 4 + 1024;
        CU(cudaMalloc(reinterpret_cast<void**>(&d_qoff), cap_qoff * 8));
      }
      ReadsPlan* d_plan = reinterpret_cast<ReadsPlan*>(d_small + 4);
      CU(reads_parse_lines(d_raw, uint32_t(n), d_nl, d_small, n_lines, fastq, eof ? 1 : 0, d_seq_len, d_is_hdr, d_seq_off,
                           d_hdr_rank, d_plan, d_qbytes, d_qoff, d_temp, temp_bytes, st));
      CU(cudaMemcpyAsync(h_plan, d_plan, sizeof(ReadsPlan), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      const ReadsPlan plan = *h_plan;
      const uint64_t nq = plan.n_records;
      if (nq) {
        Workspace::grow_dev(ws->d_qwords, ws->d_qwords_cap, size_t(packed_words(ix->alphabet, nq, plan.seq_bytes)));
        Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * 8);
        Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, size_t(nq) + 2);
        CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, st));
        {
          ProfScope p(2, r.device, st);
          CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, ws->d_qwords, ws->d_flag, st));
        }
        {
          ProfScope p(0, r.device, st);
          SearchVariant v = g_variant;
          v.avg_len = uint32_t(std::min<uint64_t>(plan.seq_bytes / nq, 1u << 30));
          CU(launch_search(r.view, ws->d_qwords, d_qoff, nq, locate ? OUT_SP_CNT_U32 : OUT_COUNT_U64, ws->d_out, ws->d_defer,
                           v, r.sm_count, st));
        }
        CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, st));
        if (!locate) {
          out.counts.resize(out.n_reads + nq);
          CU(cudaMemcpyAsync(out.counts.data() + out.n_reads, ws->d_out, nq * 8, cudaMemcpyDeviceToHost, st));
          g_prof.d2h += nq * 8;
          CU(cudaStreamSynchronize(st));
        } else {
          Workspace::grow_dev(ws->d_hit_off, ws->d_hit_off_cap, size_t(nq) + 1);
          uint64_t n_hits = locate_chunk_count(r, ws, nq, ws->d_hit_off, st);
          out.hit_off.resize(out.n_reads + nq + 1);
          uint64_t* dst_off = out.hit_off.data() + out.n_reads;
          CU(cudaMemcpyAsync(dst_off, ws->d_hit_off, (nq + 1) * 8, cudaMemcpyDeviceToHost, st));
          if (n_hits) {
            uint64_t* d_hits = locate_chunk_walk(r, ws, nq, n_hits, flags, ws->d_hit_off, st);
            if (out.n_hits + n_hits > out.hits_cap) {
              uint64_t cap = std::max<uint64_t>(out.n_hits + n_hits, out.hits_cap * 2);
              void* np = realloc(out.hits, cap * sizeof(awry_hit));
              if (!np) {
                cudaFreeAsync(d_hits, st);
                fail(AWRY_ERR_NOMEM, "out of host memory for %llu hits", (unsigned long long)cap);
              }
              out.hits = static_cast<awry_hit*>(np);
              out.hits_cap = cap;
            }
            CU(cudaMemcpyAsync(out.hits + out.n_hits, d_hits, n_hits * 16, cudaMemcpyDeviceToHost, st));
            cudaFreeAsync(d_hits, st);
            g_prof.d2h += n_hits * 16;
          }
          CU(cudaStreamSynchronize(st));
          g_prof.d2h += (nq + 1) * 8;
          for (uint64_t i = 0; i <= nq; i++) dst_off[i] += out.n_hits;
          out.n_hits += n_hits;
        }
        if (*ws->h_flag != ~0ull)
          fail(AWRY_ERR_INVALID_QUERY,
               "read %llu of %s is empty or contains a sentinel ('$'/'#'): the reference panics on it "
               "(fm_index.rs:406, bwt.rs:127)",
               (unsigned long long)(out.n_reads + *ws->h_flag), path);
        out.n_reads += nq;
        out.n_bases += plan.seq_bytes;
      }
      tail_len = n - plan.consumed;
      tail_src = base + plan.consumed;
      prev_slot = si;
      if (eof) {
        for (uint64_t i = 0; i < tail_len; i++) {
          uint8_t ch = tail_src[i];
          if (ch != '\n' && ch != '\r' && ch != ' ' && ch != '\t')
            fail(AWRY_ERR_FORMAT, "%s ends with a truncated %s record", path, fastq ? "FASTQ" : "FASTA");
        }
        break;
      }
      if (tail_len > CARRY)
        fail(AWRY_ERR_UNSUPPORTED, "%s holds a record larger than %llu bytes (AWRY_B200_READS_CHUNK)", path,
             (unsigned long long)CARRY);
    }
  } catch (...) {
    cleanup();
    free(out.hits);
    out.hits = nullptr;
    throw;
  }
  cleanup();
}

uint32_t ascii_to_dsym(const awry_index* ix, uint8_t ch, bool* sentinel) {
  if (ch >= 'a' && ch <= 'z') ch = uint8_t(ch - 'a' + 'A');
  *sentinel = (ch == '$' || ch == '#');
  if (ix->alphabet == AWRY_NUCLEOTIDE) {
    switch (ch) {
      case 'A': return 0;
      case 'C': return 1;
      case 'G': return 2;
      case 'T':
      case 'U': return 3;
      case '$':
      case '#': return DNA_SENTINEL;
      default: return DNA_N;
    }
  }
  static const char L[23] = "$ACDEFGHIKLMNPQRSTVWXY";
  if (*sentinel) return 0;
  for (int i = 1; i < 22; i++)
    if (i != 20 && L[i] == char(ch)) return uint32_t(i);
  return 20;
}

const awry_index* need(const awry_index* ix) {
  if (!ix || ix->reps.empty()) fail(AWRY_ERR_INVALID_ARG, "null or empty index handle");
  return ix;
}

}  // namespace

// ------------------------------------------------------------------ extern "C"

extern "C" {

const char* awry_last_error(void) { return g_err; }
const char* awry_version(void) { return "awry_b200 0.1.0 (sm_100a)"; }

int awry_index_load(const char* path, const int* devices, int n_dev, awry_index** out) {
  return guarded([&] {
    if (!path || !out) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) fail(AWRY_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
    // large stdio buffer: the file is read once, sequentially
    setvbuf(f, nullptr, _IOFBF, 8u << 20);
    char label[11];
    if (fread(label, 1, 11, f) != 11 || memcmp(label, "AWRY-Index\n", 11) != 0)
      fail(AWRY_ERR_FORMAT, "file did not start with the expected label; probably not an fm index file");
    uint64_t hdr[4];
    if (fread(hdr, 8, 4, f) != 4) fail(AWRY_ERR_FORMAT, "truncated header");
    auto ix = std::make_unique<awry_index>();
    ix->version = hdr[0];  // read, not validated (fm_index_file.rs:185-186)
    ix->sa_ratio = hdr[1];
    ix->bwt_len = hdr[2];
    if (hdr[3] > 1) fail(AWRY_ERR_FORMAT, "invalid symbol alphabet id %llu", (unsigned long long)hdr[3]);
    ix->alphabet = int(hdr[3]);
    check_header(ix.get());
    auto devs = pick_devices(devices, n_dev);
    Source src;
    src.f = f;
    try {
      make_replicas(ix.get(), devs, src, true);
    } catch (...) {
      awry_index_free(ix.release());
      throw;
    }
    *out = ix.release();
  });
}

int awry_index_from_parts(const awry_parts* p, const int* devices, int n_dev, awry_index** out) {
  return guarded([&] {
    if (!p || !out || !p->blocks || !p->prefix_sums || !p->sa_words) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    auto ix = std::make_unique<awry_index>();
    ix->version = p->version ? p->version : 1;
    ix->sa_ratio = p->sa_ratio;
    ix->bwt_len = p->bwt_len;
    if (p->alphabet > 1) fail(AWRY_ERR_INVALID_ARG, "invalid alphabet id %u", p->alphabet);
    ix->alphabet = int(p->alphabet);
    ix->kmer_len_file = p->kmer_len;
    check_header(ix.get());
    auto devs = pick_devices(devices, n_dev);
    for (uint64_t i = 0; i < p->n_sequences; i++) {
      ix->seq_starts.push_back(p->seq_starts ? p->seq_starts[i] : 0);
      ix->headers.emplace_back(p->headers && p->headers[i] ? p->headers[i] : "");
    }
    const uint64_t n_ref_blocks = (ix->bwt_len + 255) / 256;
    Source src;  // same read order as the file: blocks, prefix sums, SA words
    src.segs.emplace_back(reinterpret_cast<const uint8_t*>(p->blocks),
                          size_t(n_ref_blocks) * (ix->alphabet == AWRY_NUCLEOTIDE ? 160 : 352));
    src.segs.emplace_back(reinterpret_cast<const uint8_t*>(p->prefix_sums), size_t(ix->card + 1) * 8);
    src.segs.emplace_back(reinterpret_cast<const uint8_t*>(p->sa_words), size_t(ix->n_sa_words) * 8);
    try {
      make_replicas(ix.get(), devs, src, false);
    } catch (...) {
      awry_index_free(ix.release());
      throw;
    }
    *out = ix.release();
  });
}

uint64_t awry_parts_num_blocks(uint64_t bwt_len) { return (bwt_len + 255) / 256; }
uint64_t awry_parts_block_words(uint32_t alphabet) { return alphabet == AWRY_NUCLEOTIDE ? 20 : 44; }
uint64_t awry_parts_sa_words(uint64_t bwt_len, uint64_t sa_ratio) {
  return bwt_len >= 2 && sa_ratio ? sa_word_len(bwt_len, sa_ratio) : 0;
}

int awry_build_parts(uint32_t alphabet, const uint8_t* text, uint64_t n, uint64_t sa_ratio, int device,
                     uint64_t* blocks, uint64_t* prefix_sums, uint64_t* sa_words, double* phase_seconds) {
  return guarded([&] {
    if (!text || !blocks || !prefix_sums || !sa_words) fail(AWRY_ERR_INVALID_ARG, "null argument");
    if (alphabet > 1) fail(AWRY_ERR_INVALID_ARG, "invalid alphabet id %u", alphabet);
    if (n + 1 >= (1ull << 32) - 256) fail(AWRY_ERR_UNSUPPORTED, "text of %llu symbols: 32-bit row pointers", (unsigned long long)n);
    pick_devices(&device, 1);
    std::string err;
    if (build_parts(int(alphabet), text, n, sa_ratio ? sa_ratio : 8, device, blocks, prefix_sums, sa_words,
                    phase_seconds, err) != 0)
      fail(AWRY_ERR_CUDA, "index construction failed: %s", err.c_str());
  });
}

int awry_read_sequence_file(const char* path, uint32_t alphabet, uint8_t** text, uint64_t* n_text, uint64_t** starts,
                            uint64_t* n_records) {
  return guarded([&] {
    if (!path || !text || !n_text || !starts || !n_records) fail(AWRY_ERR_INVALID_ARG, "null argument");
    if (alphabet > 1) fail(AWRY_ERR_INVALID_ARG, "invalid alphabet id %u", alphabet);
    *text = nullptr;
    *starts = nullptr;
    *n_text = *n_records = 0;
    TextBuf t;
    std::string err;
    std::vector<uint64_t> st;
    std::vector<std::string> hd;
    if (read_sequence_file(path, alphabet == 0 ? 'N' : 'X', t, st, hd, err) != 0) fail(AWRY_ERR_IO, "%s", err.c_str());
    uint8_t* tb = static_cast<uint8_t*>(malloc(t.size() + 1));
    uint64_t* sb = static_cast<uint64_t*>(malloc(st.size() * 8 + 8));
    if (!tb || !sb) {
      free(tb);
      free(sb);
      fail(AWRY_ERR_NOMEM, "out of host memory");
    }
    memcpy(tb, t.data(), t.size());
    memcpy(sb, st.data(), st.size() * 8);
    *text = tb;
    *n_text = t.size();
    *starts = sb;
    *n_records = st.size();
  });
}

int awry_index_build(const awry_build_args* a, const int* devices, int n_dev, awry_index** out) {
  return guarded([&] {
    if (out) *out = nullptr;
    if (!a || !a->input_file_src) fail(AWRY_ERR_INVALID_ARG, "null argument");
    if (!a->output_file_src && !out) fail(AWRY_ERR_INVALID_ARG, "neither an output file nor an index handle requested");
    if (a->alphabet > 1) fail(AWRY_ERR_INVALID_ARG, "invalid alphabet id %u", a->alphabet);
    const int alphabet = int(a->alphabet);
    const uint64_t ratio = a->suffix_array_compression_ratio ? a->suffix_array_compression_ratio : 8;  // fm_index.rs:122
    const uint32_t k = a->lookup_table_kmer_len ? a->lookup_table_kmer_len : (alphabet == 0 ? 10u : 4u);
    const int card = alphabet == 0 ? 6 : 22;
    if (k > 255 || ipow(uint64_t(card - 2), k) > (1ull << 34)) fail(AWRY_ERR_UNSUPPORTED, "k-mer table of length %u too large", k);
    // the sequence file first: I/O and format errors do not need a device
    const bool verbose = getenv("AWRY_B200_BUILD_VERBOSE") != nullptr;  // phase times on stderr
    auto tick = [t = std::chrono::steady_clock::now(), verbose](const char* what) mutable {
      auto now = std::chrono::steady_clock::now();
      if (verbose) fprintf(stderr, "[awry_index_build] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
      t = now;
    };
    TextBuf text;
    std::string err;
    std::vector<uint64_t> starts;
    std::vector<std::string> headers;
    if (read_sequence_file(a->input_file_src, alphabet == 0 ? 'N' : 'X', text, starts, headers, err) != 0)
      fail(AWRY_ERR_IO, "%s", err.c_str());
    tick("read sequence file");
    const uint64_t n = text.size(), bwt_len = n + 1;
    if (bwt_len >= (1ull << 32) - 256) fail(AWRY_ERR_UNSUPPORTED, "text of %llu symbols: 32-bit row pointers", (unsigned long long)n);
    int dev0 = a->device;
    std::vector<int> devs = out ? pick_devices(devices ? devices : &dev0, devices ? n_dev : 1) : pick_devices(&dev0, 1);
    // construction on devs[0]; the reference-layout arrays stay on the device and are re-laid out there
    // (no 3.5 GB round trip through host memory); they are copied out only to write a file
    std::vector<uint64_t> prefix(size_t(card) + 1);
    DeviceParts dp;
    struct DpGuard {
      DeviceParts& d;
      ~DpGuard() { d.release(); }
    } dp_guard{dp};
    double ph[8] = {0};
    if (build_parts(alphabet, text.data(), n, ratio, devs[0], nullptr, prefix.data(), nullptr, ph, err, &dp) != 0)
      fail(AWRY_ERR_CUDA, "index construction failed: %s", err.c_str());
    text.reset();
    if (verbose)
      fprintf(stderr, "[awry_index_build]   ingest %.3f keys %.3f sort %.3f ties %.3f bwt %.3f milestones %.3f sa-pack %.3f\n",
              ph[0], ph[1], ph[2], ph[3], ph[4], ph[5], ph[6]);
    tick("suffix sort + BWT on device");
    auto ixp = std::make_unique<awry_index>();
    ixp->version = 1;
    ixp->sa_ratio = ratio;
    ixp->bwt_len = bwt_len;
    ixp->alphabet = alphabet;
    ixp->kmer_len_file = out ? k : 0;  // a file-only build needs no seed table
    check_header(ixp.get());
    ixp->seq_starts = starts;
    ixp->headers = headers;
    {
      Source src;
      src.dev_blocks = dp.d_blocks;
      src.dev_sa = dp.d_sa_words;
      src.segs.emplace_back(reinterpret_cast<const uint8_t*>(prefix.data()), size_t(card + 1) * 8);
      g_skip_accelerators = !out;
      try {
        make_replicas(ixp.get(), devs, src, false);
      } catch (...) {
        g_skip_accelerators = false;
        awry_index_free(ixp.release());
        throw;
      }
      g_skip_accelerators = false;
    }
    awry_index* ix = ixp.release();
    std::unique_ptr<awry_index, void (*)(awry_index*)> holder(ix, awry_index_free);
    tick("device layout + accelerators");
    if (a->output_file_src) {  // FmIndex::save (fm_index_file.rs:42-106)
      FILE* f = fopen(a->output_file_src, "wb");
      if (!f) fail(AWRY_ERR_IO, "cannot create %s: %s", a->output_file_src, strerror(errno));
      std::unique_ptr<FILE, int (*)(FILE*)> closer(f, fclose);
      setvbuf(f, nullptr, _IOFBF, 8u << 20);
      bool ok = true;
      auto put = [&](const void* p, size_t nbytes) {
        if (ok && nbytes && fwrite(p, 1, nbytes, f) != nbytes) ok = false;
      };
      put("AWRY-Index\n", 11);  // fm_index_file.rs:18,47
      uint64_t hdr[4] = {1, ratio, bwt_len, uint64_t(alphabet)};
      put(hdr, sizeof hdr);
      // device arrays -> file through a pinned double buffer (D2H of chunk i+1 overlaps fwrite of chunk i)
      auto put_device = [&](const uint64_t* d_src, uint64_t n_words) {
        DeviceGuard dg(dp.device);
        const uint64_t CH = (64u << 20) / 8;
        uint64_t* hb[2] = {nullptr, nullptr};
        cudaStream_t cs = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        try {
          CU(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
          for (int i = 0; i < 2; i++) {
            CU(cudaHostAlloc(reinterpret_cast<void**>(&hb[i]), CH * 8, cudaHostAllocDefault));
            CU(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
          }
          uint64_t n_chunks = (n_words + CH - 1) / CH;
          auto issue = [&](uint64_t c) {
            uint64_t w0 = c * CH, nw = std::min(CH, n_words - w0);
            CU(cudaMemcpyAsync(hb[c & 1], d_src + w0, nw * 8, cudaMemcpyDeviceToHost, cs));
            CU(cudaEventRecord(ev[c & 1], cs));
          };
          if (n_chunks) issue(0);
          for (uint64_t c = 0; c < n_chunks; c++) {
            CU(cudaEventSynchronize(ev[c & 1]));
            if (c + 1 < n_chunks) issue(c + 1);
            put(hb[c & 1], std::min(CH, n_words - c * CH) * 8);
          }
        } catch (...) {
          for (int i = 0; i < 2; i++) {
            cudaFreeHost(hb[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
          }
          if (cs) cudaStreamDestroy(cs);
          throw;
        }
        for (int i = 0; i < 2; i++) {
          cudaFreeHost(hb[i]);
          cudaEventDestroy(ev[i]);
        }
        cudaStreamDestroy(cs);
      };
      put_device(dp.d_blocks, dp.n_block_words);
      put(prefix.data(), prefix.size() * 8);
      put_device(dp.d_sa_words, sa_word_len(bwt_len, ratio));
      uint8_t kb = uint8_t(k);
      put(&kb, 1);
      {
        Replica& r = *ix->reps[0];
        DeviceGuard dg(r.device);
        const uint64_t n_entries = ipow(uint64_t(card - 2), k), CH = 1u << 22;
        ulonglong2* d_buf = nullptr;
        CU(cudaMalloc(reinterpret_cast<void**>(&d_buf), std::min(n_entries, CH) * 16));
        std::vector<uint64_t> h_buf(2 * std::min(n_entries, CH));
        IndexView v = r.view;
        v.kmer_len = 0;
        for (uint64_t first = 0; first < n_entries && ok; first += CH) {
          uint64_t cnt = std::min(CH, n_entries - first);
          cudaError_t e = launch_ref_table(v, first, cnt, k, d_buf, nullptr);
          if (e == cudaSuccess) e = cudaMemcpy(h_buf.data(), d_buf, cnt * 16, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) {
            cudaFree(d_buf);
            fail(AWRY_ERR_CUDA, "k-mer table kernel failed: %s", cudaGetErrorString(e));
          }
          put(h_buf.data(), cnt * 16);
        }
        cudaFree(d_buf);
      }
      uint64_t n_seqs = starts.size();  // sequence_index.rs:144-152
      put(&n_seqs, 8);
      for (uint64_t i = 0; i < n_seqs; i++) {
        uint64_t hl = headers[i].size();
        put(&starts[i], 8);
        put(&hl, 8);
        put(headers[i].data(), hl);
      }
      if (fflush(f) != 0) ok = false;
      if (!ok) fail(AWRY_ERR_IO, "write to %s failed", a->output_file_src);
      tick("write .awry file");
    }
    if (out) *out = holder.release();
  });
}

int awry_build_index_file(const awry_build_args* a) {
  if (a && !a->output_file_src) {
    return guarded([&] { fail(AWRY_ERR_INVALID_ARG, "output_file_src is null"); });
  }
  return awry_index_build(a, nullptr, 0, nullptr);
}

void awry_index_free(awry_index* ix) {
  if (!ix) return;
  for (auto& rp : ix->reps) {
    Replica& r = *rp;
    cudaSetDevice(r.device);
    for (Workspace* w : r.all_ws) {
      w->destroy();
      delete w;
    }
    cudaFree(r.d_blocks);
    cudaFree(r.d_sa);
    cudaFree(r.d_table);
    cudaFree(r.d_pair);
    cudaFree(r.d_full_sa);
    cudaFree(r.d_seq_starts);
    cudaFree(r.d_async_flag);
  }
  delete ix;
}

int awry_index_info(const awry_index* ix, awry_info* info) {
  return guarded([&] {
    need(ix);
    if (!info) fail(AWRY_ERR_INVALID_ARG, "null info");
    memset(info, 0, sizeof *info);
    info->version = ix->version;
    info->sa_ratio = ix->sa_ratio;
    info->bwt_len = ix->bwt_len;
    info->alphabet = uint32_t(ix->alphabet);
    info->kmer_len = ix->kmer_len_file;
    info->n_prefix_sums = uint32_t(ix->card + 1);
    info->n_devices = uint32_t(ix->reps.size());
    memcpy(info->prefix_sums, ix->prefix_sums, sizeof info->prefix_sums);
    info->n_sequences = ix->seq_starts.size();
    info->device_bytes_blocks = ix->reps[0]->bytes_blocks;
    info->device_bytes_sa = ix->reps[0]->bytes_sa;
    info->device_bytes_table = ix->reps[0]->bytes_table;
    info->device_bytes_pair = ix->reps[0]->bytes_pair;
    info->device_bytes_full_sa = ix->reps[0]->bytes_full_sa;
    for (size_t i = 0; i < ix->reps.size() && i < 16; i++) info->devices[i] = ix->reps[i]->device;
  });
}

int awry_index_sequence_header(const awry_index* ix, uint64_t seq_idx, const char** header, uint64_t* header_len) {
  return guarded([&] {
    need(ix);
    if (seq_idx >= ix->headers.size()) fail(AWRY_ERR_INVALID_ARG, "sequence index %llu out of range", (unsigned long long)seq_idx);
    if (header) *header = ix->headers[seq_idx].c_str();
    if (header_len) *header_len = ix->headers[seq_idx].size();
  });
}

int awry_count_batch(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq, uint64_t* counts) {
  return guarded([&] {
    need(ix);
    if (nq == 0) return;
    if (!qbytes || !qoff || !counts) fail(AWRY_ERR_INVALID_ARG, "null argument");
    validate_offsets(qoff, nq);
    for_each_replica_range(ix, qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
      search_on_replica(ix, *ix->reps[ri], qbytes, qoff, lo, hi, OUT_COUNT_U64, counts);
    });
  });
}

int awry_search_batch(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq, awry_range* ranges) {
  return guarded([&] {
    need(ix);
    if (nq == 0) return;
    if (!qbytes || !qoff || !ranges) fail(AWRY_ERR_INVALID_ARG, "null argument");
    validate_offsets(qoff, nq);
    for_each_replica_range(ix, qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
      search_on_replica(ix, *ix->reps[ri], qbytes, qoff, lo, hi, OUT_RANGE_U64, ranges);
    });
  });
}

int awry_locate_batch(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq, uint32_t flags,
                      uint64_t* hit_off, awry_hit** hits, uint64_t* n_hits) {
  return guarded([&] {
    need(ix);
    if (!hit_off || !hits || !n_hits) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *hits = nullptr;
    *n_hits = 0;
    hit_off[0] = 0;
    if (nq == 0) return;
    if (!qbytes || !qoff) fail(AWRY_ERR_INVALID_ARG, "null argument");
    validate_offsets(qoff, nq);
    size_t nr = ix->reps.size();
    std::vector<LocatePart> parts(nr);
    std::vector<std::pair<uint64_t, uint64_t>> ranges(nr, {0, 0});
    try {
      for_each_replica_range(ix, qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
        ranges[ri] = {lo, hi};
        locate_on_replica(ix, *ix->reps[ri], qbytes, qoff, lo, hi, flags, parts[ri]);
      });
    } catch (...) {
      for (auto& p : parts) free(p.hits);
      throw;
    }
    // concatenate in range order (results of the reference's order-preserving collect)
    uint64_t total = 0;
    for (auto& p : parts) total += p.n_hits;
    awry_hit* all = nullptr;
    if (nr == 1) {
      all = parts[0].hits;
      parts[0].hits = nullptr;
    } else if (total) {
      all = static_cast<awry_hit*>(malloc(total * sizeof(awry_hit)));
      if (!all) {
        for (auto& p : parts) free(p.hits);
        fail(AWRY_ERR_NOMEM, "out of host memory for %llu hits", (unsigned long long)total);
      }
    }
    uint64_t base = 0;
    for (size_t ri = 0; ri < nr; ri++) {
      auto [lo, hi] = ranges[ri];
      if (hi > lo)
        for (uint64_t i = 0; i <= hi - lo; i++) hit_off[lo + i] = base + parts[ri].hit_off[i];
      if (nr > 1 && parts[ri].n_hits) memcpy(all + base, parts[ri].hits, parts[ri].n_hits * sizeof(awry_hit));
      base += parts[ri].n_hits;
      if (nr > 1) free(parts[ri].hits);
    }
    hit_off[nq] = total;
    *hits = all;
    *n_hits = total;
  });
}

void awry_hits_free(awry_hit* hits) { free(hits); }

int awry_locate_batch_into(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq,
                           uint32_t flags, uint64_t* hit_off, awry_hit* hits, uint64_t capacity, uint64_t* n_hits) {
  return guarded([&] {
    need(ix);
    if (!hit_off || !n_hits || (!hits && capacity)) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *n_hits = 0;
    hit_off[0] = 0;
    if (nq == 0) return;
    if (!qbytes || !qoff) fail(AWRY_ERR_INVALID_ARG, "null argument");
    validate_offsets(qoff, nq);
    if (ix->reps.size() == 1) {  // hits and offsets land straight in the caller's (ideally pinned) buffers
      LocatePart part;
      part.hits = hits;
      part.ext_cap = capacity;
      part.ext_off = hit_off;  // marks the part as caller-owned even when capacity is 0
      locate_on_replica(ix, *ix->reps[0], qbytes, qoff, 0, nq, flags, part);
      *n_hits = part.n_hits;
      if (part.n_hits > capacity)
        fail(AWRY_ERR_CAPACITY, "hit buffer holds %llu entries, %llu needed", (unsigned long long)capacity,
             (unsigned long long)part.n_hits);
      return;
    }
    awry_hit* tmp = nullptr;
    uint64_t n = 0;
    int rc = awry_locate_batch(ix, qbytes, qoff, nq, flags, hit_off, &tmp, &n);
    if (rc != AWRY_OK) fail(rc, "%s", g_err);
    *n_hits = n;
    if (n > capacity) {
      free(tmp);
      fail(AWRY_ERR_CAPACITY, "hit buffer holds %llu entries, %llu needed", (unsigned long long)capacity,
           (unsigned long long)n);
    }
    if (n) memcpy(hits, tmp, n * sizeof(awry_hit));
    free(tmp);
  });
}

int awry_count_reads_file(const awry_index* ix, const char* path, uint64_t** counts, uint64_t* n_reads) {
  return guarded([&] {
    need(ix);
    if (!path || !counts || !n_reads) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *counts = nullptr;
    *n_reads = 0;
    ReadsOut out;
    run_reads_file(ix, path, false, 0, out);
    uint64_t* buf = static_cast<uint64_t*>(malloc(std::max<size_t>(8, out.n_reads * 8)));
    if (!buf) fail(AWRY_ERR_NOMEM, "out of host memory for %llu counts", (unsigned long long)out.n_reads);
    if (out.n_reads) memcpy(buf, out.counts.data(), out.n_reads * 8);
    *counts = buf;
    *n_reads = out.n_reads;
  });
}

int awry_locate_reads_file(const awry_index* ix, const char* path, uint32_t flags, uint64_t** hit_off, awry_hit** hits,
                           uint64_t* n_reads, uint64_t* n_hits) {
  return guarded([&] {
    need(ix);
    if (!path || !hit_off || !hits || !n_reads || !n_hits) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *hit_off = nullptr;
    *hits = nullptr;
    *n_reads = *n_hits = 0;
    ReadsOut out;
    run_reads_file(ix, path, true, flags, out);
    uint64_t* off = static_cast<uint64_t*>(malloc((out.n_reads + 1) * 8));
    if (!off) {
      free(out.hits);
      fail(AWRY_ERR_NOMEM, "out of host memory for %llu offsets", (unsigned long long)out.n_reads);
    }
    memcpy(off, out.hit_off.data(), (out.n_reads + 1) * 8);
    *hit_off = off;
    *hits = out.hits;
    *n_reads = out.n_reads;
    *n_hits = out.n_hits;
  });
}

void awry_buffer_free(void* p) { free(p); }

int awry_initial_range(const awry_index* ix, uint8_t ascii_symbol, awry_range* out) {
  return guarded([&] {
    need(ix);
    if (!out) fail(AWRY_ERR_INVALID_ARG, "null out");
    bool sent = false;
    uint32_t d = ascii_to_dsym(ix, ascii_symbol, &sent);
    // SearchRange::new accepts the sentinel too: [C[0], C[1]-1] (search.rs:43-48)
    const IndexView& v = ix->reps[0]->view;
    if (sent) {
      out->start_ptr = ix->prefix_sums[0];
      out->end_ptr = ix->prefix_sums[1] - 1;
    } else {
      out->start_ptr = v.c_lo[d];
      out->end_ptr = v.c_hi[d];
    }
  });
}

int awry_update_range(const awry_index* ix, awry_range range, uint8_t ascii_symbol, awry_range* out) {
  return guarded([&] {
    need(ix);
    if (!out) fail(AWRY_ERR_INVALID_ARG, "null out");
    bool sent = false;
    uint32_t d = ascii_to_dsym(ix, ascii_symbol, &sent);
    if (sent) fail(AWRY_ERR_INVALID_QUERY, "cannot extend a range with the sentinel (reference panics, bwt.rs:127)");
    if (range.start_ptr == 0 || range.start_ptr > ix->bwt_len || range.end_ptr >= ix->bwt_len)
      fail(AWRY_ERR_INVALID_ARG, "range [%llu,%llu] outside the BWT (reference: unchecked out-of-bounds read)",
           (unsigned long long)range.start_ptr, (unsigned long long)range.end_ptr);
    Replica& r = *ix->reps[0];
    DeviceGuard dg(r.device);
    Workspace* ws = r.acquire();
    try {
      Workspace::grow_dev(ws->d_out, ws->d_out_cap, 16);
      CU(launch_single_update(r.view, uint32_t(range.start_ptr), uint32_t(range.end_ptr), d,
                              reinterpret_cast<uint32_t*>(ws->d_out), ws->st));
      uint32_t res[2];
      CU(cudaMemcpyAsync(res, ws->d_out, 8, cudaMemcpyDeviceToHost, ws->st));
      CU(cudaStreamSynchronize(ws->st));
      out->start_ptr = res[0];
      out->end_ptr = res[1];
    } catch (...) {
      r.release(ws);
      throw;
    }
    r.release(ws);
  });
}

int awry_backstep(const awry_index* ix, uint64_t row, uint64_t* out) {
  return guarded([&] {
    need(ix);
    if (!out) fail(AWRY_ERR_INVALID_ARG, "null out");
    if (row >= ix->bwt_len) fail(AWRY_ERR_INVALID_ARG, "row %llu outside the BWT", (unsigned long long)row);
    Replica& r = *ix->reps[0];
    DeviceGuard dg(r.device);
    Workspace* ws = r.acquire();
    try {
      Workspace::grow_dev(ws->d_out, ws->d_out_cap, 16);
      CU(launch_single_backstep(r.view, uint32_t(row), reinterpret_cast<uint32_t*>(ws->d_out), ws->st));
      uint32_t res = 0;
      CU(cudaMemcpyAsync(&res, ws->d_out, 4, cudaMemcpyDeviceToHost, ws->st));
      CU(cudaStreamSynchronize(ws->st));
      *out = res;
    } catch (...) {
      r.release(ws);
      throw;
    }
    r.release(ws);
  });
}

// ---- device-resident entry points ----

int awry_count_device(const awry_index* ix, int replica, const uint8_t* d_qbytes, const uint64_t* d_qoff, uint64_t nq,
                      uint64_t* d_counts, void* cuda_stream) {
  return guarded([&] {
    need(ix);
    if (replica < 0 || size_t(replica) >= ix->reps.size()) fail(AWRY_ERR_INVALID_ARG, "replica out of range");
    if (nq == 0) return;
    if (!d_qbytes || !d_qoff || !d_counts) fail(AWRY_ERR_INVALID_ARG, "null argument");
    Replica& r = *ix->reps[size_t(replica)];
    DeviceGuard dg(r.device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    // total query bytes bound the packed size; read the last offset (one 8-byte D2H)
    uint64_t ends[2] = {0, 0};
    CU(cudaMemcpyAsync(&ends[0], d_qoff, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&ends[1], d_qoff + nq, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ends[1] < ends[0]) fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone");
    const int sh = packed_unit_shift(ix->alphabet);
    uint64_t words = 4 * (nq + ((ends[1] >> sh) - (ends[0] >> sh)) + 2) + 32;
    uint64_t* d_qwords = nullptr;  // packed queries, then the deferred-query list
    CU(cudaMallocAsync(reinterpret_cast<void**>(&d_qwords), words * 8 + (nq + 2) * 4, st));
    uint32_t* d_defer = reinterpret_cast<uint32_t*>(d_qwords + words);
    {
      ProfScope p(2, r.device, st);
      CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, d_qwords - 4 * (ends[0] >> sh), r.d_async_flag, st));
    }
    {
      ProfScope p(0, r.device, st);
      SearchVariant v = g_variant;
      v.avg_len = uint32_t(std::min<uint64_t>((ends[1] - ends[0]) / nq, 1u << 30));
      CU(launch_search(r.view, d_qwords - 4 * (ends[0] >> sh), d_qoff, nq, OUT_COUNT_U64, d_counts, d_defer, v, r.sm_count, st));
    }
    CU(cudaFreeAsync(d_qwords, st));
  });
}

int awry_locate_device(const awry_index* ix, int replica, const uint8_t* d_qbytes, const uint64_t* d_qoff, uint64_t nq,
                       uint32_t flags, uint64_t* d_hit_off, awry_hit** d_hits, uint64_t* n_hits, void* cuda_stream) {
  return guarded([&] {
    need(ix);
    if (replica < 0 || size_t(replica) >= ix->reps.size()) fail(AWRY_ERR_INVALID_ARG, "replica out of range");
    if (!d_hit_off || !d_hits || !n_hits) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *d_hits = nullptr;
    *n_hits = 0;
    if (nq == 0) return;
    if (!d_qbytes || !d_qoff) fail(AWRY_ERR_INVALID_ARG, "null argument");
    Replica& r = *ix->reps[size_t(replica)];
    DeviceGuard dg(r.device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint64_t ends[2] = {0, 0};
    CU(cudaMemcpyAsync(&ends[0], d_qoff, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&ends[1], d_qoff + nq, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ends[1] < ends[0]) fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone");
    const int sh = packed_unit_shift(ix->alphabet);
    Workspace* ws = r.acquire();
    try {
      uint64_t words = 4 * (nq + ((ends[1] >> sh) - (ends[0] >> sh)) + 2) + 32;
      Workspace::grow_dev(ws->d_qwords, ws->d_qwords_cap, size_t(words));
      Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * 8);
      Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, size_t(nq) + 2);
      {
        ProfScope p(2, r.device, st);
        CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, ws->d_qwords - 4 * (ends[0] >> sh), r.d_async_flag, st));
      }
      {
        ProfScope p(0, r.device, st);
        SearchVariant v = g_variant;
        v.avg_len = uint32_t(std::min<uint64_t>((ends[1] - ends[0]) / nq, 1u << 30));
        CU(launch_search(r.view, ws->d_qwords - 4 * (ends[0] >> sh), d_qoff, nq, OUT_SP_CNT_U32, ws->d_out, ws->d_defer, v, r.sm_count, st));
      }
      uint64_t n = 0;
      uint64_t* h = locate_chunk_device(ix, r, ws, nq, flags, d_hit_off, &n, st);
      CU(cudaStreamSynchronize(st));
      *d_hits = reinterpret_cast<awry_hit*>(h);
      *n_hits = n;
    } catch (...) {
      cudaStreamSynchronize(st);
      r.release(ws);
      throw;
    }
    r.release(ws);
  });
}

int awry_device_free(const awry_index* ix, int replica, void* d_ptr) {
  return guarded([&] {
    need(ix);
    if (replica < 0 || size_t(replica) >= ix->reps.size()) fail(AWRY_ERR_INVALID_ARG, "replica out of range");
    if (!d_ptr) return;
    DeviceGuard dg(ix->reps[size_t(replica)]->device);
    CU(cudaFreeAsync(d_ptr, nullptr));  // came from the stream-ordered pool
  });
}

int awry_device_check(const awry_index* ix, int replica, void* cuda_stream) {
  return guarded([&] {
    need(ix);
    if (replica < 0 || size_t(replica) >= ix->reps.size()) fail(AWRY_ERR_INVALID_ARG, "replica out of range");
    Replica& r = *ix->reps[size_t(replica)];
    DeviceGuard dg(r.device);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    unsigned long long flag = ~0ull;
    CU(cudaMemcpyAsync(&flag, r.d_async_flag, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemsetAsync(r.d_async_flag, 0xff, 8, st));
    CU(cudaStreamSynchronize(st));
    if (flag != ~0ull)
      fail(AWRY_ERR_INVALID_QUERY, "query %llu of a device batch is empty or contains a sentinel", flag);
  });
}

// ---- instrumentation ----

int awry_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.enabled = on != 0;
  return AWRY_OK;
}
int awry_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  prof_collect_locked();
  for (int i = 0; i < 3; i++) g_prof.n[i] = 0, g_prof.ms[i] = 0;
  g_prof.h2d = 0;
  g_prof.d2h = 0;
  kernel_launch_count_reset();
  return AWRY_OK;
}
int awry_profile_get(awry_profile* out) {
  if (!out) return AWRY_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(g_prof.mu);
  prof_collect_locked();
  out->launches = kernel_launch_count();
  out->search_launches = g_prof.n[0];
  out->search_ms = g_prof.ms[0];
  out->walk_launches = g_prof.n[1];
  out->walk_ms = g_prof.ms[1];
  out->pack_launches = g_prof.n[2];
  out->pack_ms = g_prof.ms[2];
  out->h2d_bytes = g_prof.h2d;
  out->d2h_bytes = g_prof.d2h;
  return AWRY_OK;
}

int awry_bench_random_gather(int device, uint64_t footprint_bytes, uint32_t granule, uint32_t lanes, uint64_t n_reads,
                             int iters, double* reads_per_s, double* gb_per_s) {
  return guarded([&] {
    if (!reads_per_s || !gb_per_s || footprint_bytes < granule || n_reads == 0) fail(AWRY_ERR_INVALID_ARG, "bad argument");
    pick_devices(&device, 1);
    DeviceGuard dg(device);
    cudaError_t e = run_random_gather(footprint_bytes, granule, lanes, n_reads, iters, reads_per_s, gb_per_s);
    if (e == cudaErrorInvalidValue) fail(AWRY_ERR_INVALID_ARG, "unsupported granule/lanes combination %u/%u", granule, lanes);
    CU(e);
  });
}

int awry_set_host_pack(int mode) {
  return guarded([&] {
    if (mode < -1 || mode > 1) fail(AWRY_ERR_INVALID_ARG, "host pack mode must be -1 (auto), 0 (off) or 1 (on)");
    g_host_pack = mode;
  });
}

int awry_host_pack_dna(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t* exceptions, uint64_t exc_cap, uint64_t* n_exc) {
  return guarded([&] {
    if (!src || !dst || !n_exc) fail(AWRY_ERR_INVALID_ARG, "null argument");
    std::vector<uint64_t> exc;
    host_pack_dna(src, size_t(n), dst, exc, 0);
    *n_exc = exc.size();
    if (exceptions)
      for (size_t i = 0; i < exc.size() && i < exc_cap; i++) exceptions[i] = exc[i];
  });
}

int awry_set_locate_variant(int variant) {
  return guarded([&] {
    if (variant != 0 && variant != 1) fail(AWRY_ERR_INVALID_ARG, "locate variant must be 0 (default) or 1 (LF-walk)");
    g_locate_variant = variant;
  });
}

int awry_set_search_variant(int lanes_per_query, int threads_per_block, int blocks_per_sm) {
  if (lanes_per_query != 0 && lanes_per_query != -1 && lanes_per_query != 1 && lanes_per_query != 2 &&
      lanes_per_query != 4 && lanes_per_query != 8)
    return AWRY_ERR_INVALID_ARG;
  g_variant.lanes = lanes_per_query;
  g_variant.tpb = threads_per_block;
  g_variant.blocks_per_sm = blocks_per_sm;
  return AWRY_OK;
}

}  // extern "C"
