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
                    bool probe_link = false, bool copy_flag = true) {
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
    trace_mark("  host pack done, bytes", (long long)nbytes);
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
    gpu_mark(ws->st, "h2d done", (long long)c.q0);
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
  gpu_mark(ws->st, "packed on device", (long long)c.q0);
  {
    ProfScope p(0, r.device, ws->st);
    SearchVariant v = g_variant;
    v.avg_len = uint32_t(std::min<uint64_t>(nbytes / std::max<uint64_t>(1, nq), 1u << 30));
    CU(launch_search(r.view, ws->d_qwords - 4 * (c.b0 >> packed_unit_shift(ix->alphabet)), ws->d_qoff, nq, mode, ws->d_out, ws->d_defer, v, r.sm_count, ws->st));
  }
  gpu_mark(ws->st, "searched", (long long)c.q0);
  // (the locate pipeline copies the flag together with the hit total, after the scan: one hand-over between
  // the compute and copy engines per chunk instead of two)
  if (copy_flag) CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, ws->st));
  trace_mark("  search enqueued, queries", (long long)nq);
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

void locate_on_replica(const awry_index* ix, Replica& r, const uint8_t* qbytes, const uint64_t* qoff,
                       uint64_t q_lo, uint64_t q_hi, uint32_t flags, LocatePart& part) {
  const bool ext = part.ext_cap != 0 || part.ext_off != nullptr;
  if (!part.ext_off) part.hit_off.assign(q_hi - q_lo + 1, 0);
  uint64_t* off_base = part.ext_off ? part.ext_off : part.hit_off.data();
  if (q_lo >= q_hi) return;
  DeviceGuard dg(r.device);
  const bool src_pinned = is_pinned(qbytes) && is_pinned(qoff);
  auto chunks = make_chunks(qoff, q_lo, q_hi, locate_chunk_q(), std::min(chunk_max_bytes(), locate_chunk_bytes()));
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
    enqueue_search(ix, r, ws, qbytes, qoff, c, OUT_SP_CNT_U32, src_pinned, true, false, false);
    mark("A searched", i);
    Workspace::grow_dev(ws->d_hit_off, ws->d_hit_off_cap, size_t(nq) + 1);
    locate_chunk_scan(ws, nq, ws->d_hit_off, ws->st);
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

}  // namespace host
}  // namespace awry

extern "C" {

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
