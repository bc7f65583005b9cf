// batch.cu -- the batched count / locate pipelines of libawry_b200 and their entry points:
// parallel_count / parallel_locate (rayon map over queries, /root/reference/src/fm_index.rs:455-487) as
// chunked host -> device -> host pipelines per replica, plus the device-resident variants.
#include "host.hpp"

using namespace awry;
using namespace awry::host;

namespace awry {
namespace host {

// ------------------------------------------------------------------ batched pipelines

// AWRY_B200_TRACE=1: host-side stage times of the batched pipelines on stderr (scripts/locate_trace.py)
static bool trace_on() {
  static const bool on = getenv("AWRY_B200_TRACE") != nullptr;
  return on;
}
static void trace_mark(const char* what, long long i = -1) {
  if (!trace_on()) return;
  static const auto t_origin = std::chrono::steady_clock::now();
  fprintf(stderr, "[trace] %10.3f ms  %s %lld\n",
          std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_origin).count(), what, i);
}
// device-side timeline of a traced call: events recorded on the chunk streams, printed relative to the first
struct GpuTrace {
  struct Ev {
    cudaEvent_t e;
    const char* what;
    long long i;
  };
  std::vector<Ev> evs;
  void mark(cudaStream_t st, const char* what, long long i) {
    if (!trace_on()) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    evs.push_back(Ev{e, what, i});
  }
  void dump() {
    if (evs.empty()) return;
    for (auto& v : evs) cudaEventSynchronize(v.e);
    for (auto& v : evs) {
      float ms = 0;
      cudaEventElapsedTime(&ms, evs[0].e, v.e);
      fprintf(stderr, "[gpu]   %10.3f ms  %s %lld\n", ms, v.what, v.i);
    }
    for (auto& v : evs) cudaEventDestroy(v.e);
    evs.clear();
  }
};
static thread_local GpuTrace* g_gpu_trace = nullptr;
static void gpu_mark(cudaStream_t st, const char* what, long long i) {
  if (g_gpu_trace) g_gpu_trace->mark(st, what, i);
}

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

// `ramp_from` > 0: the first chunk holds about that many bytes and every following one up to 1.4 x its
// predecessor, until max_bytes is reached (see shaped_chunks).
std::vector<Chunk> make_chunks(const uint64_t* qoff, uint64_t q_lo, uint64_t q_hi, uint64_t max_q, uint64_t max_bytes,
                               uint64_t ramp_from = 0) {
  std::vector<Chunk> out;
  uint64_t q = q_lo;
  double cap = ramp_from ? double(std::min(ramp_from, max_bytes)) : double(max_bytes);
  while (q < q_hi) {
    const uint64_t cap_bytes = uint64_t(cap);
    uint64_t hi = std::min(q_hi, q + max_q);
    // largest hi with qoff[hi] - qoff[q] <= cap_bytes (at least one query)
    if (qoff[hi] - qoff[q] > cap_bytes) {
      uint64_t lo2 = q + 1, hi2 = hi;
      while (lo2 < hi2) {
        uint64_t mid = (lo2 + hi2 + 1) / 2;
        if (qoff[mid] - qoff[q] <= cap_bytes)
          lo2 = mid;
        else
          hi2 = mid - 1;
      }
      hi = lo2;
    }
    out.push_back(Chunk{q, hi, qoff[q], qoff[hi]});
    q = hi;
    cap = std::min(double(max_bytes), cap * 1.4);
  }
  return out;
}

// Pipeline fill and drain: the device idles while the host prepares the first chunk, and the host idles while the
// device works on the last one.  ASCII input (the host packs every chunk before its copy): the first chunk is split
// into 1/4 + 3/4 and the last into 1/2 + 1/4 + 1/4 (by queries).  A finer ramp -- 16 MB first, every chunk 1.4 x its
// predecessor -- was measured and LOSES there (21.0-21.5 against 19.9 ms per 10 M reads,
// profiles/r02_s6_count_e2e_trace.log): packing and copying chunk i + 1 one after the other takes longer than
// searching chunk i, so the device waits a little at every step of the ramp instead of once, and chunks of 0.1-0.4 M
// reads run the persistent search kernel at 2.0-2.4 ms per M reads instead of 1.8.  Pre-packed input has no host
// pass; there the ramp wins (19.0-19.3 -> 18.5 ms) and is used (`ramp`).
std::vector<Chunk> shaped_chunks(const uint64_t* qoff, uint64_t q_lo, uint64_t q_hi, uint64_t max_q, uint64_t max_bytes,
                                 bool ramp) {
  const uint64_t total = qoff[q_hi] >= qoff[q_lo] ? qoff[q_hi] - qoff[q_lo] : 0;
  const uint64_t RAMP_FROM = 16u << 20, MIN_SPLIT = 64u << 20;  // (only chunks worth splitting are split)
  std::vector<Chunk> in = make_chunks(qoff, q_lo, q_hi, max_q, max_bytes, ramp && total >= 3 * RAMP_FROM ? RAMP_FROM : 0);
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
  if (!ramp && in.front().b1 - in.front().b0 >= MIN_SPLIT)
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
  const int hp = g_host_pack.load(std::memory_order_relaxed);
  if (hp >= 0) return hp == 1 && host_pack_supported();
  static const bool on = [] {
    if (const char* e = getenv("AWRY_B200_HOST_PACK")) return e[0] != '0' && host_pack_supported();
    return host_pack_supported() && host_pool_threads() >= 4;
  }();
  return on;
}

// Where a batch's queries come from: ASCII bytes (the reference's `&str`s), or 2-bit codes the caller holds
// already (awry_*_batch_packed2: crumb of byte position p at bits 2*(p%4) of crumbs[p/4], exceptions sorted
// by position) -- then the host packer has nothing to do and a quarter of the bytes cross PCIe.
struct QuerySource {
  const uint8_t* qbytes = nullptr;
  const uint8_t* crumbs = nullptr;
  const uint64_t* exc = nullptr;  // ((position) << 8) | byte, ascending
  uint64_t n_exc = 0;
  const uint64_t* qoff = nullptr;
  bool pinned = false;  // bytes / crumbs and offsets are page-locked: the copy engine reads them in place
};

// The chunker binary-searches the offsets and sizes every buffer from a chunk's byte range, so the ends of
// every chunk are checked here (a non-monotone array can make b1 < b0 or a chunk absurdly large); inside a
// chunk each query is checked against [b0, b1] on the device before it is touched (pack kernels).
void validate_chunks(const std::vector<Chunk>& chunks, uint64_t max_bytes) {
  for (const Chunk& c : chunks) {
    if (c.b1 < c.b0) fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone (queries %llu..%llu)",
                          (unsigned long long)c.q0, (unsigned long long)c.q1);
    // one query may exceed the chunk budget; several may not
    if (c.b1 - c.b0 > max_bytes && c.q1 - c.q0 > 1)
      fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone (queries %llu..%llu)", (unsigned long long)c.q0,
           (unsigned long long)c.q1);
    if (c.b1 - c.b0 >= (1ull << 32)) fail(AWRY_ERR_INVALID_ARG, "query %llu is longer than 4 GiB", (unsigned long long)c.q0);
  }
}

// Sequencer reads are all the same length: then a chunk's offsets are first + i * len and the device writes them
// itself (8 bytes per read that do not cross the link: 80 MB of the 455 MB a 10 M x 150-bp batch of packed reads
// uploads).  Checked exactly, on the pool, per chunk; AWRY_B200_UNIFORM_OFFSETS=0 always copies.
bool offsets_are_uniform(const uint64_t* off, uint64_t nq, uint64_t& len) {
  static const bool on = [] {
    const char* e = getenv("AWRY_B200_UNIFORM_OFFSETS");
    return !(e && e[0] == '0');
  }();
  if (!on || nq < 4096) return false;
  len = off[1] - off[0];
  if (off[1] < off[0] || len == 0 || len >= (1ull << 32) || off[nq] - off[0] != len * nq) return false;
  const uint64_t BLK = 1u << 16, nb = (nq + BLK - 1) / BLK, first = off[0], l = len;
  std::atomic<bool> ok{true};
  host_parallel_blocks(size_t(nb), [&](size_t b) {
    if (!ok.load(std::memory_order_relaxed)) return;
    const uint64_t lo = b * BLK, hi = std::min(nq, lo + BLK);
    uint64_t bad = 0;
    for (uint64_t i = lo; i <= hi; i++) bad |= off[i] ^ (first + i * l);
    if (bad) ok.store(false, std::memory_order_relaxed);
  });
  return ok.load();
}

// Uploads one chunk of queries and runs prepass + search on ws->st.  The search result lands
// in ws->d_out in the requested mode.
// `may_pack`: the caller's say on host packing for this chunk (PackBalance)
// `d_out_override`: where the search result goes instead of ws->d_out (a call-level array the chunks fill).
void enqueue_search(const awry_index* ix, Replica& r, PackBalance& bal, Workspace* ws, const QuerySource& qs,
                    const Chunk& c, SearchOut mode, bool may_pack = true, bool probe_link = false, bool copy_flag = true,
                    void* d_out_override = nullptr) {
  const uint64_t nq = c.q1 - c.q0, nbytes = c.b1 - c.b0;
  const size_t out_elem = mode == OUT_COUNT_U64 ? 8 : mode == OUT_RANGE_U64 ? 16 : sp_cnt_bytes(r.view);
  const int ush = packed_unit_shift(ix->alphabet);
  Workspace::grow_dev(ws->d_qbytes, ws->d_qbytes_cap, size_t(nbytes) + 16);
  Workspace::grow_dev(ws->d_qoff, ws->d_qoff_cap, size_t(nq) + 1);
  Workspace::grow_dev(ws->d_qwords, ws->d_qwords_cap, size_t(packed_words(ix->alphabet, nq, nbytes)));
  Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * out_elem);
  Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, defer_words(nq));
  const uint64_t* src_o = qs.qoff + c.q0;
  uint64_t* const d_words = ws->d_qwords - 4 * (c.b0 >> ush);
  ws->link_probe_bytes = 0;
  uint64_t uni_len = 0;
  const bool uniform = offsets_are_uniform(src_o, nq, uni_len);
  auto upload_offsets = [&] {
    if (uniform) {
      CU(launch_fill_offsets(ws->d_qoff, c.b0, uni_len, nq, ws->st));
    } else {
      CU(cudaMemcpyAsync(ws->d_qoff, src_o, (nq + 1) * 8, cudaMemcpyHostToDevice, ws->st));
      g_prof.h2d += (nq + 1) * 8;
    }
  };
  auto stage_offsets = [&] {  // pageable offsets: staged through pinned memory
    if (qs.pinned || uniform) return;
    Workspace::grow_host(ws->h_qoff, ws->h_qoff_cap, size_t(nq) + 1);
    parallel_memcpy(ws->h_qoff, src_o, (nq + 1) * 8);
    src_o = ws->h_qoff;
  };
  if (qs.crumbs) {
    // pre-packed by the caller: bytes [b0/4, ceil(b1/4)) of the crumb array, exceptions of [b0, b1)
    const uint64_t cb0 = c.b0 >> 2, cb1 = (c.b1 + 3) >> 2;
    const size_t pbytes = size_t(cb1 - cb0);
    const uint8_t* src_c = qs.crumbs + cb0;
    const uint64_t* e_lo = std::lower_bound(qs.exc, qs.exc + qs.n_exc, c.b0 << 8);
    const uint64_t* e_hi = std::lower_bound(e_lo, qs.exc + qs.n_exc, c.b1 << 8);
    const size_t n_exc = size_t(e_hi - e_lo);
    stage_offsets();
    if (!qs.pinned) {
      Workspace::grow_host(ws->h_qbytes, ws->h_qbytes_cap, pbytes + 64);
      parallel_memcpy(ws->h_qbytes, src_c, pbytes);
      src_c = ws->h_qbytes;
    }
    if (pbytes) CU(cudaMemcpyAsync(ws->d_qbytes, src_c, pbytes, cudaMemcpyHostToDevice, ws->st));
    upload_offsets();
    if (n_exc) {
      Workspace::grow_dev(ws->d_exc, ws->d_exc_cap, n_exc);
      const uint64_t* src_e = e_lo;
      if (!qs.pinned) {
        Workspace::grow_host(ws->h_exc, ws->h_exc_cap, n_exc);
        memcpy(ws->h_exc, e_lo, n_exc * 8);
        src_e = ws->h_exc;
      }
      CU(cudaMemcpyAsync(ws->d_exc, src_e, n_exc * 8, cudaMemcpyHostToDevice, ws->st));
    }
    g_prof.h2d += pbytes + n_exc * 8;
    gpu_mark(ws->st, "h2d done", (long long)c.q0);
    CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, ws->st));
    {
      ProfScope p(2, r.device, ws->st);
      CU(launch_pack2(reinterpret_cast<const uint32_t*>(ws->d_qbytes), cb0 << 2, ws->d_qoff, nq, d_words, ws->d_exc, n_exc,
                      0, c.b0, c.b1, ws->d_flag, ws->st));
    }
  } else {
    const uint8_t* src_b = qs.qbytes + c.b0;
    // Nucleotide chunks are packed to 2 bits per base by the host cores before the copy (a quarter of
    // the PCIe bytes; works the same for pageable and pinned caller memory).  Chunks with many bytes
    // outside ACGT go up as ASCII.
    bool packed = false;
    if (may_pack && ix->alphabet == AWRY_NUCLEOTIDE && host_pack_enabled() && nbytes >= 4096) {
      Workspace::grow_host(ws->h_qbytes, ws->h_qbytes_cap, size_t(nbytes) / 4 + 64);
      auto t0 = std::chrono::steady_clock::now();
      int sharers = 1;
      packed = host_pack_dna(src_b, size_t(nbytes), ws->h_qbytes, ws->exc_tmp, 64, 0, &sharers);
      if (packed)
        bal.note_host(double(nbytes), std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), sharers);
      trace_mark("  host pack done, bytes", (long long)nbytes);
    }
    if (packed) {
      const size_t n_exc = ws->exc_tmp.size();
      const size_t pbytes = (size_t(nbytes) + 3) / 4;
      stage_offsets();
      CU(cudaMemcpyAsync(ws->d_qbytes, ws->h_qbytes, pbytes + 8, cudaMemcpyHostToDevice, ws->st));
      upload_offsets();
      if (n_exc) {
        Workspace::grow_host(ws->h_exc, ws->h_exc_cap, n_exc);
        Workspace::grow_dev(ws->d_exc, ws->d_exc_cap, n_exc);
        memcpy(ws->h_exc, ws->exc_tmp.data(), n_exc * 8);
        CU(cudaMemcpyAsync(ws->d_exc, ws->h_exc, n_exc * 8, cudaMemcpyHostToDevice, ws->st));
      }
      g_prof.h2d += pbytes + 8 + n_exc * 8;
      gpu_mark(ws->st, "h2d done", (long long)c.q0);
      CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, ws->st));
      {
        ProfScope p(2, r.device, ws->st);
        CU(launch_pack2(reinterpret_cast<const uint32_t*>(ws->d_qbytes), c.b0, ws->d_qoff, nq, d_words, ws->d_exc, n_exc,
                        c.b0, c.b0, c.b1, ws->d_flag, ws->st));
      }
    } else {
      if (!qs.pinned) {
        Workspace::grow_host(ws->h_qbytes, ws->h_qbytes_cap, size_t(nbytes) + 16);
        parallel_memcpy(ws->h_qbytes, src_b, nbytes);
        src_b = ws->h_qbytes;
      }
      stage_offsets();
      const bool probe = probe_link && qs.pinned && nbytes >= (8u << 20);
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
      upload_offsets();
      g_prof.h2d += nbytes;
      CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, ws->st));
      // offsets stay absolute: the kernels subtract the chunk's byte base
      {
        ProfScope p(2, r.device, ws->st);
        CU(launch_pack(ix->alphabet, ws->d_qbytes - c.b0, ws->d_qoff, nq, d_words, c.b0, c.b1, ws->d_flag, ws->st));
      }
    }
  }
  gpu_mark(ws->st, "packed on device", (long long)c.q0);
  {
    ProfScope p(0, r.device, ws->st);
    SearchVariant v = current_variant();
    v.avg_len = uint32_t(std::min<uint64_t>(nbytes / std::max<uint64_t>(1, nq), 1u << 30));
    v.b_lo = c.b0;
    v.b_hi = c.b1;
    // locate: when pass 2 will be the gather from the unsampled array, queries finished in the text are stored as
    // their text position (decided here, once, and remembered in the workspace for locate_chunk_walk)
    v.locate_positions = mode == OUT_SP_CNT_U32 && locate_positions_ok(r.view, v);
    ws->sp_cnt_positions = v.locate_positions;
    ws->gpu_probe_bytes = 0;
    const bool time_kernel = qs.pinned && !qs.crumbs && nbytes >= (8u << 20);  // PackBalance: what the GPU consumes
    if (time_kernel) {
      if (!ws->ev_s0) {
        CU(cudaEventCreate(&ws->ev_s0));
        CU(cudaEventCreate(&ws->ev_s1));
      }
      CU(cudaEventRecord(ws->ev_s0, ws->st));
    }
    CU(launch_search(r.view, d_words, ws->d_qoff, nq, mode, d_out_override ? d_out_override : ws->d_out, ws->d_defer, v,
                     r.sm_count, ws->st));
    if (time_kernel) {
      CU(cudaEventRecord(ws->ev_s1, ws->st));
      ws->gpu_probe_bytes = nbytes;
    }
  }
  gpu_mark(ws->st, "searched", (long long)c.q0);
  // (the locate pipeline copies the flag together with the hit total, after the scan: one hand-over between
  // the compute and copy engines per chunk instead of two)
  if (copy_flag) CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, ws->st));
  trace_mark("  search enqueued, queries", (long long)nq);
}

[[noreturn]] void fail_bad_query(unsigned long long code, uint64_t q_base) {
  const unsigned long long q = q_base + (code >> 1);
  if ((code & 1) == BAD_OFFSETS)
    fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone at query %llu (its bytes lie outside the batch)", q);
  fail(AWRY_ERR_INVALID_QUERY,
       "query %llu is empty or contains a sentinel ('$'/'#'): the reference panics on it "
       "(fm_index.rs:406, bwt.rs:127)", q);
}

void check_flag(Workspace* ws, const Chunk& c) {
  if (*ws->h_flag != ~0ull) fail_bad_query(*ws->h_flag, c.q0);
}

// count / range search over [q_lo, q_hi) on one replica, 3-deep pipeline
void search_on_replica(const awry_index* ix, size_t ri, const QuerySource& qs, uint64_t q_lo, uint64_t q_hi,
                       SearchOut mode, void* out, size_t n_active_replicas = 1) {
  if (q_lo >= q_hi) return;
  Replica& r = *ix->reps[ri];
  PackBalance& bal = ix->balance_of(ri);
  DeviceGuard dg(r.device);
  const size_t out_elem = mode == OUT_COUNT_U64 ? 8 : 16;
  const uint64_t* qoff = qs.qoff;
  const bool src_pinned = qs.pinned;
  const bool dst_pinned = is_pinned(out);
  // pre-packed queries are a quarter of the bytes: four times the reads per chunk keep the chunk count down
  const uint64_t max_bytes = chunk_max_bytes() * (qs.crumbs ? 2 : 1);
  if (qoff[q_hi] < qoff[q_lo]) fail(AWRY_ERR_INVALID_ARG, "query offsets are not monotone (queries %llu..%llu)",
                                     (unsigned long long)q_lo, (unsigned long long)q_hi);
  auto chunks = shaped_chunks(qoff, q_lo, q_hi, CHUNK_MAX_Q, max_bytes, qs.crumbs != nullptr);
  validate_chunks(chunks, max_bytes);
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
        bal.note_link(double(ws[s]->link_probe_bytes), double(ms) * 1e-3);
      ws[s]->link_probe_bytes = 0;
    }
    if (ws[s]->gpu_probe_bytes) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ws[s]->ev_s0, ws[s]->ev_s1) == cudaSuccess)
        bal.note_gpu(double(ws[s]->gpu_probe_bytes), double(ms) * 1e-3);
      ws[s]->gpu_probe_bytes = 0;
    }
    check_flag(ws[s], c);
    if (!dst_pinned)
      parallel_memcpy(static_cast<char*>(out) + c.q0 * out_elem, ws[s]->h_out, (c.q1 - c.q0) * out_elem);
    pending[s] = -1;
  };
  GpuTrace gpu_trace;  // AWRY_B200_TRACE=1: device-side timeline of the chunk pipeline on stderr
  g_gpu_trace = trace_on() ? &gpu_trace : nullptr;
  uint64_t bytes_total = 0, bytes_packed = 0;  // of the chunks enqueued so far (pinned sources only)
  const bool balanced = src_pinned && !qs.crumbs && ix->alphabet == AWRY_NUCLEOTIDE && host_pack_enabled();
  PackBalance::Plan plan{1.0, false, false};
  if (balanced) plan = bal.plan(n_active_replicas);
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
          // a chunk goes up as it is as soon as the raw share is behind plan: raw chunks sit early in the batch,
          // where the host packs the following chunks while they cross the link, not at its end, where the host
          // would idle behind them (measured effect: none on a host whose memory system is the limit -- 15.4-15.9 ms
          // per 10 M reads for every share from 0.7 to 1.0, profiles/r02_s10_e2e_share_sweep.log)
          may_pack = !(double(bytes_total - bytes_packed) < (1.0 - plan.share) * double(bytes_total + (c.b1 - c.b0)));
        }
        bytes_total += c.b1 - c.b0;
        if (may_pack) bytes_packed += c.b1 - c.b0;
      }
      enqueue_search(ix, r, bal, ws[s], qs, c, mode, may_pack, probe);
      size_t bytes = (c.q1 - c.q0) * out_elem;
      void* dst = static_cast<char*>(out) + c.q0 * out_elem;
      if (!dst_pinned) {
        Workspace::grow_host(ws[s]->h_out, ws[s]->h_out_cap, bytes);
        dst = ws[s]->h_out;
      }
      CU(cudaMemcpyAsync(dst, ws[s]->d_out, bytes, cudaMemcpyDeviceToHost, ws[s]->st));
      gpu_mark(ws[s]->st, "results copied", (long long)c.q0);
      g_prof.d2h += bytes;
      CU(cudaEventRecord(ws[s]->done, ws[s]->st));
      pending[s] = int(i);
    }
    for (int s = 0; s < DEPTH; s++) finish(s);
    g_gpu_trace = nullptr;
    gpu_trace.dump();
  } catch (...) {
    g_gpu_trace = nullptr;
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
void locate_chunk_scan(Replica& r, Workspace* ws, uint64_t nq, uint64_t* d_hit_off, cudaStream_t st) {
  const void* d_sp_cnt = ws->d_out;
  size_t temp = 0;
  CU(scan_hit_offsets(r.view, d_sp_cnt, nq, d_hit_off, nullptr, temp, st));
  Workspace::grow_dev(reinterpret_cast<uint8_t*&>(ws->d_temp), ws->d_temp_cap, temp + 16);
  CU(scan_hit_offsets(r.view, d_sp_cnt, nq, d_hit_off, ws->d_temp, temp, st));
}

// device-side two-pass locate of a chunk whose queries were already searched (ws->d_out holds
// (sp, count) per query).  Step 1: CSR offsets + hit total (one synchronisation).
uint64_t locate_chunk_count(Replica& r, Workspace* ws, uint64_t nq, uint64_t* d_hit_off, cudaStream_t st) {
  locate_chunk_scan(r, ws, nq, d_hit_off, st);
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
  const int lv = g_locate_variant.load(std::memory_order_relaxed);
  if (lv != 0 && !ws->sp_cnt_positions) view.full_sa = nullptr;  // 1, 2: walk (unless pass 1 already wrote positions)
  if (lv == 1) view.walk_blocks = nullptr;                        // 1: to the file's row samples
  const void* d_sp_cnt = ws->d_out;
  uint64_t* d_hits = nullptr;
  CU(cudaMallocAsync(reinterpret_cast<void**>(&d_hits), n_hits * 16 + 16, st));  // pool: no driver round trip
  trace_mark("  hits buffer allocated, hits", (long long)n_hits);
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
// device->host read (the hit total sizes pass 2), and three chunks are in flight: while the host waits for
// chunk i's total, chunk i+1 is being packed, copied and searched, and chunk i-1's hits are on their way back.
// Chunk size: every chunk pays ~0.4 ms of fixed device-side latency (a persistent search kernel that fills
// the part and drains, a dozen small dependent operations, copies slowed 2-3x by the search kernel of the
// neighbouring chunk: profiles/r01_s50_locate_gpu_timeline.log), so 1 M x 50-bp queries take 3.31 / 2.55 /
// 2.51 / 2.71 ms end to end with chunks of 256 k / 384 k / 512 k / 1 M queries
// (profiles/r01_s51_locate_chunk_sweep.log): 512 k.
constexpr uint64_t LOCATE_CHUNK_BYTES = 64u << 20;
static uint64_t locate_chunk_q() {  // AWRY_B200_LOCATE_CHUNK_Q: queries per chunk (experiments; read per call)
  if (const char* e = getenv("AWRY_B200_LOCATE_CHUNK_Q")) return std::min<uint64_t>(1u << 24, std::max<uint64_t>(1024, strtoull(e, nullptr, 10)));
  return 1u << 19;
}
static uint64_t locate_chunk_bytes() {
  if (const char* e = getenv("AWRY_B200_LOCATE_CHUNK_MB")) return std::min<uint64_t>(1024, std::max<uint64_t>(1, strtoull(e, nullptr, 10))) << 20;
  return LOCATE_CHUNK_BYTES;
}

void locate_on_replica(const awry_index* ix, size_t ri, const QuerySource& qs, uint64_t q_lo, uint64_t q_hi,
                       uint32_t flags, LocatePart& part) {
  const bool ext = part.ext_cap != 0 || part.ext_off != nullptr;
  if (!part.ext_off) part.hit_off.assign(q_hi - q_lo + 1, 0);
  uint64_t* off_base = part.ext_off ? part.ext_off : part.hit_off.data();
  if (q_lo >= q_hi) return;
  Replica& r = *ix->reps[ri];
  PackBalance& bal = ix->balance_of(ri);
  DeviceGuard dg(r.device);
  const uint64_t* qoff = qs.qoff;
  const uint64_t max_bytes = std::min(chunk_max_bytes(), locate_chunk_bytes());
  auto chunks = make_chunks(qoff, q_lo, q_hi, locate_chunk_q(), max_bytes);
  validate_chunks(chunks, max_bytes);
  constexpr int DEPTH = 3;
  struct Slot {
    Workspace* ws = nullptr;
    int chunk = -1;
    int phase = 0;  // 1 = searched + scanned (total on its way), 2 = pass 2 enqueued (results on their way)
    uint64_t base = 0;
  } slot[DEPTH];
  size_t cap = 0;
  auto mark = [&](const char* what, int i) { trace_mark(what, i); };
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
    gpu_mark(ws->st, "chunk begins", i);
    enqueue_search(ix, r, bal, ws, qs, c, OUT_SP_CNT_U32, true, false, false);
    mark("A searched", i);
    Workspace::grow_dev(ws->d_hit_off, ws->d_hit_off_cap, size_t(nq) + 1);
    locate_chunk_scan(r, ws, nq, ws->d_hit_off, ws->st);
    gpu_mark(ws->st, "scanned", i);
    CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, ws->st));
    CU(cudaMemcpyAsync(ws->h_total, ws->d_hit_off + nq, 8, cudaMemcpyDeviceToHost, ws->st));
    CU(cudaEventRecord(ws->done, ws->st));
    gpu_mark(ws->st, "total copied", i);
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
    gpu_mark(ws->st, "pass 2 begins", i);
    CU(cudaMemcpyAsync(off_base + (c.q0 - q_lo), ws->d_hit_off, nq * 8, cudaMemcpyDeviceToHost, ws->st));
    gpu_mark(ws->st, "offsets copied", i);
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
      gpu_mark(ws->st, "hits located", i);
      CU(cudaMemcpyAsync(part.hits + part.n_hits, d_hits, n_hits * 16, cudaMemcpyDeviceToHost, ws->st));
      gpu_mark(ws->st, "hits copied", i);
      cudaFreeAsync(d_hits, ws->st);
      g_prof.d2h += n_hits * 16;
    }
    CU(cudaEventRecord(ws->done, ws->st));
    mark("B end", i);
    s.phase = 2;
    part.n_hits += n_hits;  // keeps counting past the capacity so the caller learns the need
  };
  GpuTrace gpu_trace;
  g_gpu_trace = trace_on() ? &gpu_trace : nullptr;
  try {
    stage_a(0);
    for (int i = 0; i < int(chunks.size()); i++) {
      if (i + 1 < int(chunks.size())) stage_a(i + 1);
      stage_b(i);
    }
    for (auto& s : slot) stage_c(s);
    off_base[q_hi - q_lo] = part.n_hits;
    mark("done", int(chunks.size()));
    g_gpu_trace = nullptr;
    gpu_trace.dump();
  } catch (...) {
    g_gpu_trace = nullptr;
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

// parallel_locate into the caller's pinned buffers with no host round trip between the passes (the
// unsampled-SA gather, BWT order).  Per chunk: pack -> search -> scan -> the device hands the chunk its hit base
// (advance_hit_base_kernel, ordered behind the previous chunk's by an event) -> the gather writes the hits
// straight into `hits` (host memory, through its device view) and the rebased offsets into a device array
// that one copy returns.  The host enqueues chunk after chunk and waits once at the end; a workspace is
// reused when its chunk of three rounds ago has drained.  Returns the hit total (it may exceed capacity).
// cfg3 end to end: 2.2 ms -> see profiles/ (the two-sync pipeline above stays for library-owned results,
// sorted output and the LF-walk variants, which size their buffers from the total).
uint64_t locate_direct_on_replica(const awry_index* ix, size_t ri, const QuerySource& qs, uint64_t nq, uint64_t* hit_off,
                                  void* hits_dev, uint64_t capacity) {
  Replica& r = *ix->reps[ri];
  PackBalance& bal = ix->balance_of(ri);
  DeviceGuard dg(r.device);
  const uint64_t* qoff = qs.qoff;
  const uint64_t max_bytes = std::min(chunk_max_bytes(), locate_chunk_bytes());
  auto chunks = make_chunks(qoff, 0, nq, locate_chunk_q(), max_bytes);
  validate_chunks(chunks, max_bytes);
  // A call of this size is latency-bound and the device idles while the host packs the first chunk (0.4 ms for
  // 25 MB: profiles/r02_s2_locate_e2e_trace.log).  So the first chunk is cut to a quarter and goes up as it is --
  // the copy engine starts at once, the device packs it -- while the host packs what follows.
  bool first_raw = false;
  if (chunks.size() >= 1 && chunks[0].q1 - chunks[0].q0 >= (1u << 16) && qs.pinned && !qs.crumbs &&
      getenv("AWRY_B200_LOCATE_FIRST_RAW") == nullptr) {
    const Chunk c = chunks[0];
    const uint64_t cut = c.q0 + (c.q1 - c.q0) / 4;
    chunks[0] = Chunk{c.q0, cut, qoff[c.q0], qoff[cut]};
    chunks.insert(chunks.begin() + 1, Chunk{cut, c.q1, qoff[cut], qoff[c.q1]});
    first_raw = true;
  }
  const bool off_pinned = is_pinned(hit_off);
  constexpr int DEPTH = 3;
  Workspace* ws[DEPTH] = {nullptr, nullptr, nullptr};
  int pending[DEPTH] = {-1, -1, -1};
  auto finish = [&](int s) {
    if (pending[s] < 0) return;
    const Chunk& c = chunks[size_t(pending[s])];
    CU(cudaEventSynchronize(ws[s]->done));
    check_flag(ws[s], c);
    if (!off_pinned) memcpy(hit_off + c.q0, ws[s]->h_out, (c.q1 - c.q0) * 8);
    pending[s] = -1;
  };
  GpuTrace gpu_trace;
  g_gpu_trace = trace_on() ? &gpu_trace : nullptr;
  uint64_t total = 0;
  try {
    ws[0] = r.acquire();
    unsigned long long* const d_running = ws[0]->d_base + 1;
    CU(cudaMemsetAsync(d_running, 0, 8, ws[0]->st));
    CU(cudaEventRecord(ws[0]->ev_adv, ws[0]->st));  // chunk 0 waits for the counter's reset like for a predecessor
    cudaEvent_t prev_adv = ws[0]->ev_adv;
    for (size_t i = 0; i < chunks.size(); i++) {
      const int s = int(i % DEPTH);
      if (!ws[s]) ws[s] = r.acquire();
      finish(s);
      Workspace* w = ws[s];
      const Chunk& c = chunks[i];
      const uint64_t n = c.q1 - c.q0;
      gpu_mark(w->st, "chunk begins", (long long)i);
      enqueue_search(ix, r, bal, w, qs, c, OUT_SP_CNT_U32, !(first_raw && i == 0), false, false);
      Workspace::grow_dev(w->d_hit_off, w->d_hit_off_cap, size_t(n) + 1);
      Workspace::grow_dev(w->d_off_out, w->d_off_out_cap, size_t(n) + 1);
      locate_chunk_scan(r, w, n, w->d_hit_off, w->st);
      gpu_mark(w->st, "scanned", (long long)i);
      // (chunk 0 waits for the counter's reset; a slot's event is re-recorded only after the wait on its
      // earlier state has been enqueued)
      CU(launch_gather_direct(r.view, reinterpret_cast<const uint2*>(w->d_out), w->d_hit_off, n, d_running, w->d_base,
                              hits_dev, capacity, w->d_off_out, r.sm_count, prev_adv, w->ev_adv, w->st));
      prev_adv = w->ev_adv;
      gpu_mark(w->st, "hits written", (long long)i);
      uint64_t* dst = hit_off + c.q0;
      if (!off_pinned) {
        Workspace::grow_host(w->h_out, w->h_out_cap, size_t(n) * 8);
        dst = reinterpret_cast<uint64_t*>(w->h_out);
      }
      CU(cudaMemcpyAsync(dst, w->d_off_out, n * 8, cudaMemcpyDeviceToHost, w->st));
      CU(cudaMemcpyAsync(w->h_flag, w->d_flag, 8, cudaMemcpyDeviceToHost, w->st));
      if (i + 1 == chunks.size()) CU(cudaMemcpyAsync(w->h_total, d_running, 8, cudaMemcpyDeviceToHost, w->st));
      g_prof.d2h += n * 8 + 16;
      CU(cudaEventRecord(w->done, w->st));
      pending[s] = int(i);
    }
    const int last = int((chunks.size() - 1) % DEPTH);
    for (int s = 0; s < DEPTH; s++) finish((last + 1 + s) % DEPTH);  // oldest first; the last chunk carries the total
    total = *ws[last]->h_total;
    g_prof.d2h += std::min(total, capacity) * sizeof(awry_hit);
    hit_off[nq] = total;
    g_gpu_trace = nullptr;
    gpu_trace.dump();
  } catch (...) {
    g_gpu_trace = nullptr;
    for (int s = 0; s < DEPTH; s++)
      if (ws[s]) {
        cudaStreamSynchronize(ws[s]->st);
        r.release(ws[s]);
      }
    throw;
  }
  for (int s = 0; s < DEPTH; s++)
    if (ws[s]) r.release(ws[s]);
  return total;
}

// the device's view of a page-locked host buffer (cudaHostAlloc / cudaHostRegister), or nullptr
void* device_view_of_pinned(const void* p) {
  if (!p || !is_pinned(p)) return nullptr;
  void* d = nullptr;
  if (cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return d;
}

static bool direct_locate_enabled() {  // AWRY_B200_LOCATE_DIRECT=0: always the two-sync pipeline (A/B runs)
  const char* e = getenv("AWRY_B200_LOCATE_DIRECT");
  return !(e && e[0] == '0');
}

QuerySource ascii_source(const uint8_t* qbytes, const uint64_t* qoff) {
  QuerySource qs;
  qs.qbytes = qbytes;
  qs.qoff = qoff;
  qs.pinned = is_pinned(qbytes) && is_pinned(qoff);
  return qs;
}

// awry_*_batch_packed2 arguments -> QuerySource (nucleotide only; exceptions must ascend by position)
QuerySource packed2_source(const awry_index* ix, const uint8_t* crumbs, const uint64_t* qoff, uint64_t nq,
                           const uint64_t* exceptions, uint64_t n_exc) {
  if (ix->alphabet != AWRY_NUCLEOTIDE) fail(AWRY_ERR_UNSUPPORTED, "2-bit packed queries exist for the nucleotide alphabet only");
  if (!crumbs || !qoff || (n_exc && !exceptions)) fail(AWRY_ERR_INVALID_ARG, "null argument");
  for (uint64_t i = 1; i < n_exc; i++)
    if ((exceptions[i] >> 8) <= (exceptions[i - 1] >> 8))
      fail(AWRY_ERR_INVALID_ARG, "exception list is not strictly ascending by position (entry %llu)", (unsigned long long)i);
  if (n_exc && (exceptions[n_exc - 1] >> 8) >= qoff[nq] && qoff[nq] >= qoff[0])
    fail(AWRY_ERR_INVALID_ARG, "exception position %llu lies behind the last query byte",
         (unsigned long long)(exceptions[n_exc - 1] >> 8));
  QuerySource qs;
  qs.crumbs = crumbs;
  qs.exc = exceptions;
  qs.n_exc = n_exc;
  qs.qoff = qoff;
  qs.pinned = is_pinned(crumbs) && is_pinned(qoff) && (n_exc == 0 || is_pinned(exceptions));
  return qs;
}

void count_impl(const awry_index* ix, const QuerySource& qs, uint64_t nq, SearchOut mode, void* out) {
  const size_t nr = ix->reps.size();
  const size_t active = (nr == 1 || nq < 2 * nr) ? 1 : nr;  // (for_each_replica_range's own rule)
  for_each_replica_range(ix, qs.qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
    search_on_replica(ix, ri, qs, lo, hi, mode, out, active);
  });
}

void locate_owned_impl(const awry_index* ix, const QuerySource& qs, uint64_t nq, uint32_t flags, uint64_t* hit_off,
                       awry_hit** hits, uint64_t* n_hits) {
  size_t nr = ix->reps.size();
  std::vector<LocatePart> parts(nr);
  std::vector<std::pair<uint64_t, uint64_t>> ranges(nr, {0, 0});
  try {
    for_each_replica_range(ix, qs.qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
      ranges[ri] = {lo, hi};
      locate_on_replica(ix, ri, qs, lo, hi, flags, parts[ri]);
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
}

// parallel_locate over SEVERAL replicas into the caller's pinned buffers (BWT order, unsampled array).  A
// replica's hits start behind those of the replicas before it, which nobody knows until they have searched; so
// the call has two phases with one meeting of the replica threads between them.  Phase 1, all replicas at once:
// the chunk pipeline packs, uploads and searches, the (sp, count) pairs land in one array per replica, one scan
// gives its local CSR offsets and its hit total.  The totals are prefix-summed on the host.  Phase 2, all
// replicas at once: one gather per replica writes the hits STRAIGHT into the caller's buffer at its base, the
// rebased offsets go back with one copy.  No staging buffer, no concatenation pass on the host.
struct ReplicaLocateState {
  uint64_t lo = 0, hi = 0, total = 0, base = 0;
  Workspace* ws = nullptr;  // holds the call-level arrays of this replica (d_out: pairs, d_hit_off: local offsets)
};

void locate_multi_phase1(const awry_index* ix, size_t ri, const QuerySource& qs, ReplicaLocateState& st, size_t n_active) {
  (void)n_active;
  if (st.lo >= st.hi) return;
  Replica& r = *ix->reps[ri];
  PackBalance& bal = ix->balance_of(ri);
  DeviceGuard dg(r.device);
  const uint64_t n = st.hi - st.lo;
  const uint64_t max_bytes = std::min(chunk_max_bytes(), locate_chunk_bytes());
  auto chunks = make_chunks(qs.qoff, st.lo, st.hi, locate_chunk_q(), max_bytes);
  validate_chunks(chunks, max_bytes);
  st.ws = r.acquire();
  Workspace* top = st.ws;
  const size_t pair_bytes = sp_cnt_bytes(r.view);
  Workspace::grow_dev(top->d_out, top->d_out_cap, size_t(n) * pair_bytes);
  Workspace::grow_dev(top->d_hit_off, top->d_hit_off_cap, size_t(n) + 1);
  Workspace::grow_dev(top->d_off_out, top->d_off_out_cap, size_t(n) + 1);
  constexpr int DEPTH = 3;
  Workspace* ws[DEPTH] = {nullptr, nullptr, nullptr};
  int pending[DEPTH] = {-1, -1, -1};
  auto finish = [&](int s) {
    if (pending[s] < 0) return;
    CU(cudaEventSynchronize(ws[s]->done));
    check_flag(ws[s], chunks[size_t(pending[s])]);
    pending[s] = -1;
  };
  try {
    for (size_t i = 0; i < chunks.size(); i++) {
      const int s = int(i % DEPTH);
      if (!ws[s]) ws[s] = r.acquire();
      finish(s);
      const Chunk& c = chunks[i];
      enqueue_search(ix, r, bal, ws[s], qs, c, OUT_SP_CNT_U32, true, false, true,
                     top->d_out + size_t(c.q0 - st.lo) * pair_bytes);
      CU(cudaEventRecord(ws[s]->done, ws[s]->st));
      pending[s] = int(i);
    }
    for (int s = 0; s < DEPTH; s++) finish(s);  // every chunk has searched: the pairs are complete
    locate_chunk_scan(r, top, n, top->d_hit_off, top->st);
    CU(cudaMemcpyAsync(top->h_total, top->d_hit_off + n, 8, cudaMemcpyDeviceToHost, top->st));
    CU(cudaStreamSynchronize(top->st));
    st.total = *top->h_total;
  } catch (...) {
    for (int s = 0; s < DEPTH; s++)
      if (ws[s]) {
        cudaStreamSynchronize(ws[s]->st);
        r.release(ws[s]);
      }
    cudaStreamSynchronize(top->st);
    r.release(top);
    st.ws = nullptr;
    throw;
  }
  for (int s = 0; s < DEPTH; s++)
    if (ws[s]) r.release(ws[s]);
}

void locate_multi_phase2(const awry_index* ix, size_t ri, ReplicaLocateState& st, uint64_t* hit_off, void* hits_dev,
                         uint64_t capacity) {
  if (st.lo >= st.hi || !st.ws) return;
  Replica& r = *ix->reps[ri];
  DeviceGuard dg(r.device);
  Workspace* top = st.ws;
  const uint64_t n = st.hi - st.lo;
  const bool off_pinned = is_pinned(hit_off);
  try {
    unsigned long long base = st.base;
    CU(cudaMemcpyAsync(top->d_base + 1, &base, 8, cudaMemcpyHostToDevice, top->st));  // (pageable source: copied at the call)
    CU(launch_gather_direct(r.view, reinterpret_cast<const uint2*>(top->d_out), top->d_hit_off, n, top->d_base + 1, top->d_base,
                            hits_dev, capacity, top->d_off_out, r.sm_count, nullptr, nullptr, top->st));
    uint64_t* dst = hit_off + st.lo;
    if (!off_pinned) {
      Workspace::grow_host(top->h_out, top->h_out_cap, size_t(n) * 8);
      dst = reinterpret_cast<uint64_t*>(top->h_out);
    }
    CU(cudaMemcpyAsync(dst, top->d_off_out, n * 8, cudaMemcpyDeviceToHost, top->st));
    g_prof.d2h += n * 8 + std::min(st.total, capacity > st.base ? capacity - st.base : 0) * sizeof(awry_hit);
    CU(cudaStreamSynchronize(top->st));
    if (!off_pinned) parallel_memcpy(hit_off + st.lo, top->h_out, n * 8);
  } catch (...) {
    cudaStreamSynchronize(top->st);
    r.release(top);
    st.ws = nullptr;
    throw;
  }
  r.release(top);
  st.ws = nullptr;
}

void locate_into_impl(const awry_index* ix, const QuerySource& qs, uint64_t nq, uint32_t flags, uint64_t* hit_off,
                      awry_hit* hits, uint64_t capacity, uint64_t* n_hits) {
  auto too_small = [&](uint64_t need) {
    fail(AWRY_ERR_CAPACITY, "hit buffer holds %llu entries, %llu needed", (unsigned long long)capacity,
         (unsigned long long)need);
  };
  if (ix->reps.size() == 1) {
    Replica& r0 = *ix->reps[0];
    void* hits_dev = capacity ? device_view_of_pinned(hits) : nullptr;
    if (hits_dev && !(flags & AWRY_LOCATE_SORTED) && r0.view.full_sa && g_locate_variant == 0 && direct_locate_enabled()) {
      *n_hits = locate_direct_on_replica(ix, 0, qs, nq, hit_off, hits_dev, capacity);
      if (*n_hits > capacity) too_small(*n_hits);
      return;
    }
    // hits and offsets land straight in the caller's (ideally pinned) buffers
    LocatePart part;
    part.hits = hits;
    part.ext_cap = capacity;
    part.ext_off = hit_off;  // marks the part as caller-owned even when capacity is 0
    locate_on_replica(ix, 0, qs, 0, nq, flags, part);
    *n_hits = part.n_hits;
    if (part.n_hits > capacity) too_small(part.n_hits);
    return;
  }
  {  // several replicas: two phases with the hits written straight into the caller's pinned buffer
    void* hits_dev = capacity ? device_view_of_pinned(hits) : nullptr;
    bool all_full_sa = true;
    for (auto& rp : ix->reps) all_full_sa = all_full_sa && rp->view.full_sa != nullptr;
    const size_t nr = ix->reps.size();
    if (hits_dev && !(flags & AWRY_LOCATE_SORTED) && all_full_sa && g_locate_variant == 0 && direct_locate_enabled() &&
        nq >= 2 * nr) {
      std::vector<ReplicaLocateState> st(nr);
      auto release_all = [&] {
        for (size_t ri = 0; ri < nr; ri++)
          if (st[ri].ws) {
            cudaSetDevice(ix->reps[ri]->device);
            cudaStreamSynchronize(st[ri].ws->st);
            ix->reps[ri]->release(st[ri].ws);
            st[ri].ws = nullptr;
          }
      };
      try {
        for_each_replica_range(ix, qs.qoff, nq, [&](size_t ri, uint64_t lo, uint64_t hi) {
          st[ri].lo = lo;
          st[ri].hi = hi;
          locate_multi_phase1(ix, ri, qs, st[ri], nr);
        });
        uint64_t total = 0;
        for (auto& x : st) {
          x.base = total;
          total += x.total;
        }
        for_each_replica_range(ix, qs.qoff, nq, [&](size_t ri, uint64_t, uint64_t) {
          locate_multi_phase2(ix, ri, st[ri], hit_off, hits_dev, capacity);
        });
        hit_off[nq] = total;
        *n_hits = total;
      } catch (...) {
        release_all();
        throw;
      }
      if (*n_hits > capacity) too_small(*n_hits);
      return;
    }
  }
  awry_hit* tmp = nullptr;
  uint64_t n = 0;
  locate_owned_impl(ix, qs, nq, flags, hit_off, &tmp, &n);
  *n_hits = n;
  if (n > capacity) {
    free(tmp);
    too_small(n);
  }
  if (n) parallel_memcpy(hits, tmp, n * sizeof(awry_hit));
  free(tmp);
}

}  // namespace host
}  // namespace awry

extern "C" {

int awry_count_batch(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq, uint64_t* counts) {
  return guarded([&] {
    need(ix);
    if (nq == 0) return;
    if (!qbytes || !qoff || !counts) fail(AWRY_ERR_INVALID_ARG, "null argument");
    count_impl(ix, ascii_source(qbytes, qoff), nq, OUT_COUNT_U64, counts);
  });
}

int awry_search_batch(const awry_index* ix, const uint8_t* qbytes, const uint64_t* qoff, uint64_t nq, awry_range* ranges) {
  return guarded([&] {
    need(ix);
    if (nq == 0) return;
    if (!qbytes || !qoff || !ranges) fail(AWRY_ERR_INVALID_ARG, "null argument");
    count_impl(ix, ascii_source(qbytes, qoff), nq, OUT_RANGE_U64, ranges);
  });
}

int awry_count_batch_packed2(const awry_index* ix, const uint8_t* crumbs, const uint64_t* qoff, uint64_t nq,
                             const uint64_t* exceptions, uint64_t n_exc, uint64_t* counts) {
  return guarded([&] {
    need(ix);
    if (nq == 0) return;
    if (!counts) fail(AWRY_ERR_INVALID_ARG, "null argument");
    count_impl(ix, packed2_source(ix, crumbs, qoff, nq, exceptions, n_exc), nq, OUT_COUNT_U64, counts);
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
    locate_owned_impl(ix, ascii_source(qbytes, qoff), nq, flags, hit_off, hits, n_hits);
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
    locate_into_impl(ix, ascii_source(qbytes, qoff), nq, flags, hit_off, hits, capacity, n_hits);
  });
}

int awry_locate_batch_packed2(const awry_index* ix, const uint8_t* crumbs, const uint64_t* qoff, uint64_t nq,
                              const uint64_t* exceptions, uint64_t n_exc, uint32_t flags, uint64_t* hit_off,
                              awry_hit* hits, uint64_t capacity, uint64_t* n_hits) {
  return guarded([&] {
    need(ix);
    if (!hit_off || !n_hits || (!hits && capacity)) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *n_hits = 0;
    hit_off[0] = 0;
    if (nq == 0) return;
    locate_into_impl(ix, packed2_source(ix, crumbs, qoff, nq, exceptions, n_exc), nq, flags, hit_off, hits, capacity, n_hits);
  });
}

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
    CU(cudaMallocAsync(reinterpret_cast<void**>(&d_qwords), words * 8 + defer_words(nq) * 4, st));
    uint32_t* d_defer = reinterpret_cast<uint32_t*>(d_qwords + words);
    {
      ProfScope p(2, r.device, st);
      CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, d_qwords - 4 * (ends[0] >> sh), ends[0], ends[1], r.d_async_flag, st));
    }
    {
      ProfScope p(0, r.device, st);
      SearchVariant v = current_variant();
      v.avg_len = uint32_t(std::min<uint64_t>((ends[1] - ends[0]) / nq, 1u << 30));
      v.b_lo = ends[0];
      v.b_hi = ends[1];
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
      Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * sp_cnt_bytes(r.view));
      Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, defer_words(nq));
      {
        ProfScope p(2, r.device, st);
        CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, ws->d_qwords - 4 * (ends[0] >> sh), ends[0], ends[1], r.d_async_flag, st));
      }
      {
        ProfScope p(0, r.device, st);
        SearchVariant v = current_variant();
        v.avg_len = uint32_t(std::min<uint64_t>((ends[1] - ends[0]) / nq, 1u << 30));
        v.b_lo = ends[0];
        v.b_hi = ends[1];
        v.locate_positions = locate_positions_ok(r.view, v);
        ws->sp_cnt_positions = v.locate_positions;
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
    if (flag != ~0ull) fail_bad_query(flag, 0);
  });
}

// ---- instrumentation ----

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

int awry_set_host_threads(int n) {
  return guarded([&] {
    if (n < 0 || n > 256) fail(AWRY_ERR_INVALID_ARG, "host thread count must be 0 (default) .. 256");
    host_set_threads(n);
  });
}

int awry_host_threads(void) { return host_pool_threads(); }

int awry_set_count_variant(int variant) {
  return guarded([&] {
    if (variant < 0 || variant > 1)
      fail(AWRY_ERR_INVALID_ARG, "count variant must be 0 (default: finish one-row intervals in the text) or 1 (backward search only)");
    g_count_variant = variant;
  });
}

int awry_set_locate_variant(int variant) {
  return guarded([&] {
    if (variant < 0 || variant > 2)
      fail(AWRY_ERR_INVALID_ARG, "locate variant must be 0 (default), 1 (LF-walk to row samples) or 2 (bounded walk)");
    g_locate_variant = variant;
  });
}

int awry_set_search_variant(int lanes_per_query, int threads_per_block, int blocks_per_sm) {
  if (lanes_per_query != 0 && lanes_per_query != -1 && lanes_per_query != 1 && lanes_per_query != 2 &&
      lanes_per_query != 4 && lanes_per_query != 8 && (lanes_per_query < 80 || lanes_per_query > 84))
    return AWRY_ERR_INVALID_ARG;
  int slots = -1;  // default choice of the pair kernel's flavour
  if (lanes_per_query >= 80) {  // 80 / 81 / 82: the pair kernel with 0 (branching refill) / 1 / 2 state-machine slots; 83 / 84: 1 / 2 slots at a lower residency
    slots = lanes_per_query - 80;
    lanes_per_query = 8;
  }
  store_variant(lanes_per_query, threads_per_block, blocks_per_sm, slots);
  return AWRY_OK;
}

}  // extern "C"
