// reads_api.cu -- host side of the streaming reads-file front-end (device side: reads.cu).
#include <sys/mman.h>
#include <zlib.h>

#include "host.hpp"

using namespace awry;
using namespace awry::host;

namespace {

// ------------------------------------------------------------------ streaming reads-file front-end
// FASTQ / FASTA file of queries -> parallel_count / parallel_locate without a host-side parser:
// a reader thread preads the file into a ring of pinned buffers, the raw bytes are uploaded and
// parsed on the device (reads.cu), and the parsed (query bytes, CSR offsets) feed the same pack /
// search / locate kernels as the *_device entry points.  A record cut by a chunk boundary is carried
// into the next chunk by the host.

// counts of a reads file: a malloc'd array that grows by doubling and is handed to the caller as it is (the
// results of a chunk come back through the workspace's pinned buffer and are copied in after the chunk's wait)
struct CountsBuf {
  uint64_t* p = nullptr;
  size_t n = 0, cap = 0;
  CountsBuf() = default;
  CountsBuf(const CountsBuf&) = delete;
  CountsBuf& operator=(const CountsBuf&) = delete;
  ~CountsBuf() { free(p); }
  void resize(size_t want) {
    if (want > cap) {
      size_t c = std::max<size_t>(std::max<size_t>(want, cap * 2), 1u << 16);
      void* q = realloc(p, c * 8);
      if (!q) fail(AWRY_ERR_NOMEM, "out of host memory for %llu counts", (unsigned long long)c);
      p = static_cast<uint64_t*>(q);
      cap = c;
    }
    n = want;
  }
  uint64_t* data() { return p; }
  uint64_t* release() {
    uint64_t* r = p;
    p = nullptr;
    n = cap = 0;
    return r;
  }
};

struct ReadsOut {
  CountsBuf counts;
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

// `ri`: the replica that does the work.  [range_lo, range_hi): a byte range of a PLAIN file that begins at a
// record start and ends at one (or at the end of the file), with the format already known (`fastq_known`) --
// how run_reads_file_multi hands every replica its own segment; the default is the whole file.
// `bad_read`: receives the segment-relative number of the read an AWRY_ERR_INVALID_QUERY is about.
void run_reads_file(const awry_index* ix, const char* path, bool locate, uint32_t flags, ReadsOut& out, size_t ri = 0,
                    uint64_t range_lo = 0, uint64_t range_hi = ~0ull, int fastq_known = -1, uint64_t* bad_read = nullptr) {
  const bool ranged = range_hi != ~0ull;
  int fd = open(path, O_RDONLY);
  if (fd < 0) fail(AWRY_ERR_IO, "cannot open %s: %s", path, strerror(errno));
  struct FdCloser {
    int fd;
    ~FdCloser() { close(fd); }
  } fdc{fd};
  struct stat sb;
  if (fstat(fd, &sb) != 0) fail(AWRY_ERR_IO, "cannot stat %s", path);
  const uint64_t fsize = ranged ? std::min<uint64_t>(range_hi, uint64_t(sb.st_size)) : uint64_t(sb.st_size);
  out.file_bytes = ranged ? fsize - std::min(fsize, range_lo) : fsize;
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
  if (ranged) {
    fastq = fastq_known;
    data_start = range_lo;
    if (data_start >= fsize) {
      if (locate) out.hit_off.assign(1, 0);
      return;
    }
  } else {
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

  Replica& r = *ix->reps[ri];
  DeviceGuard dg(r.device);
  Workspace* ws = r.acquire();
  cudaStream_t st = ws->st;
  Workspace::ReadsScratch& rs = ws->rs;

  // AWRY_B200_TRACE=1: where the wall time of the call goes (one line per call on stderr)
  static const bool trace = getenv("AWRY_B200_TRACE") != nullptr;
  using Clock = std::chrono::steady_clock;
  double t_slot = 0, t_lines = 0, t_parse = 0, t_search = 0, t_pread = 0, t_ring = 0, t_copy_out = 0;  // seconds
  auto since = [](Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); };
  const auto t_call = Clock::now();
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
  // The hits array is sized for the whole segment when the first chunk is back; its pages are mapped by a helper
  // thread (madvise(MADV_POPULATE_WRITE): contents untouched) while the pipeline runs, so that the copies into it do
  // not take a page fault per 4 KiB (160 MB of hits: 75 ms of a 160 ms call).  Joined before the array moves.
  std::thread populate;
  auto populate_done = [&] {
    if (populate.joinable()) populate.join();
  };

  auto cleanup = [&] {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    if (reader.joinable()) reader.join();
    populate_done();
    if (rs.st_in) cudaStreamSynchronize(rs.st_in);
    cudaStreamSynchronize(st);
    r.release(ws);
  };
  try {
    const uint32_t max_bytes = uint32_t(CARRY + CHUNK + 1);
    if (rs.chunk != CHUNK || rs.carry != CARRY) {
      rs.release();
      for (auto& b : rs.h_buf) CU(cudaHostAlloc(reinterpret_cast<void**>(&b), CARRY + CHUNK + 64, cudaHostAllocDefault));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_raw), max_bytes + 64));
      CU(cudaMalloc(reinterpret_cast<void**>(&rs.d_raw_b), max_bytes + 64));
      CU(cudaStreamCreateWithFlags(&rs.st_in, cudaStreamNonBlocking));
      for (auto& e : rs.ev_in) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
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
    uint8_t* const d_raw2[2] = {rs.d_raw, rs.d_raw_b};
    uint8_t* d_qbytes = rs.d_qbytes;
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
          const auto t0 = Clock::now();
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return stop || !s.filled; });
          t_ring += since(t0);
          if (stop) return;
        }
        const auto t_read = Clock::now();
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
        t_pread += since(t_read);
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
    uint64_t tail_len = 0;
    const uint8_t* tail_src = nullptr;
    int prev_slot = -1;
    // a chunk as it goes to the device: the carried tail of the previous chunk + the freshly read bytes
    struct In {
      uint8_t* base = nullptr;
      uint64_t n = 0;
      bool eof = false;
      int si = -1;
    };
    // assembles chunk c in its pinned slot and starts its upload on the input stream; the upload of chunk
    // c + 1 is started as soon as chunk c's records are split (its tail is known), so it overlaps the search
    // of chunk c.  Non-blocking mode gives up if the reader thread has not delivered the chunk yet.
    auto stage_in = [&](int c, bool blocking, In& in) -> bool {
      const int si = c % NBUF;
      Slot& s = slots[si];
      {
        const auto t0 = Clock::now();
        std::unique_lock<std::mutex> lk(mu);
        if (blocking)
          cv.wait(lk, [&] { return s.filled; });
        else if (!s.filled)
          return false;
        t_slot += since(t0);
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
        prev_slot = -1;
      }
      uint64_t n = tail_len + s.n;
      if (s.eof && n && base[n - 1] != '\n') base[n++] = '\n';
      in.base = base;
      in.n = n;
      in.eof = s.eof;
      in.si = si;
      if (n) {
        CU(cudaMemcpyAsync(d_raw2[c & 1], base, n, cudaMemcpyHostToDevice, rs.st_in));
        CU(cudaEventRecord(rs.ev_in[c & 1], rs.st_in));
        g_prof.h2d += n;
      }
      return true;
    };
    In in, nxt;
    stage_in(0, true, in);
    for (int c = 0;; c++) {
      if (in.n == 0) break;
      uint8_t* const d_raw = d_raw2[c & 1];
      uint8_t* const base = in.base;
      const uint64_t n = in.n;
      const bool eof = in.eof;
      CU(cudaStreamWaitEvent(st, rs.ev_in[c & 1], 0));
      CU(reads_find_lines(d_raw, uint32_t(n), d_nl, d_small, d_temp, temp_bytes, st));
      CU(cudaMemcpyAsync(h_n_lines, d_small, 4, cudaMemcpyDeviceToHost, st));
      auto t_sync = Clock::now();
      CU(cudaStreamSynchronize(st));
      t_lines += since(t_sync);
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
      t_sync = Clock::now();
      CU(cudaStreamSynchronize(st));
      t_parse += since(t_sync);
      const ReadsPlan plan = *h_plan;
      const uint64_t nq = plan.n_records;
      tail_len = n - plan.consumed;
      tail_src = base + plan.consumed;
      prev_slot = in.si;
      bool have_next = false;
      if (eof) {
        for (uint64_t i = 0; i < tail_len; i++) {
          uint8_t ch = tail_src[i];
          if (ch != '\n' && ch != '\r' && ch != ' ' && ch != '\t')
            fail(AWRY_ERR_FORMAT, "%s ends with a truncated %s record", path, fastq ? "FASTQ" : "FASTA");
        }
      } else {
        if (tail_len > CARRY)
          fail(AWRY_ERR_UNSUPPORTED, "%s holds a record larger than %llu bytes (AWRY_B200_READS_CHUNK)", path,
               (unsigned long long)CARRY);
        have_next = stage_in(c + 1, false, nxt);  // its upload overlaps the search below
      }
      if (nq) {
        Workspace::grow_dev(ws->d_qwords, ws->d_qwords_cap, size_t(packed_words(ix->alphabet, nq, plan.seq_bytes)));
        Workspace::grow_dev(ws->d_out, ws->d_out_cap, size_t(nq) * (locate ? sp_cnt_bytes(r.view) : 8));
        Workspace::grow_dev(ws->d_defer, ws->d_defer_cap, defer_words(nq));
        CU(cudaMemsetAsync(ws->d_flag, 0xff, 8, st));
        {
          ProfScope p(2, r.device, st);
          // offsets come from the library's own record splitter: [0, seq_bytes]
          CU(launch_pack(ix->alphabet, d_qbytes, d_qoff, nq, ws->d_qwords, 0, plan.seq_bytes, ws->d_flag, st));
        }
        {
          ProfScope p(0, r.device, st);
          SearchVariant v = current_variant();
          v.avg_len = uint32_t(std::min<uint64_t>(plan.seq_bytes / nq, 1u << 30));
          v.b_lo = 0;
          v.b_hi = plan.seq_bytes;
          v.locate_positions = locate && locate_positions_ok(r.view, v);
          ws->sp_cnt_positions = v.locate_positions;
          CU(launch_search(r.view, ws->d_qwords, d_qoff, nq, locate ? OUT_SP_CNT_U32 : OUT_COUNT_U64, ws->d_out, ws->d_defer,
                           v, r.sm_count, st));
        }
        CU(cudaMemcpyAsync(ws->h_flag, ws->d_flag, 8, cudaMemcpyDeviceToHost, st));
        if (!locate) {
          if (out.counts.cap == 0 && plan.consumed) {  // first chunk: size the array for the whole segment
            const double per_read = double(plan.consumed) / double(nq);
            out.counts.resize(size_t(double(out.file_bytes) / per_read * 1.02) + nq + 1024);
          }
          out.counts.resize(out.n_reads + nq);
          Workspace::grow_host(ws->h_out, ws->h_out_cap, size_t(nq) * 8);
          CU(cudaMemcpyAsync(ws->h_out, ws->d_out, nq * 8, cudaMemcpyDeviceToHost, st));
          g_prof.d2h += nq * 8;
          t_sync = Clock::now();
          CU(cudaStreamSynchronize(st));
          t_search += since(t_sync);
          memcpy(out.counts.data() + out.n_reads, ws->h_out, nq * 8);
        } else {
          Workspace::grow_dev(ws->d_hit_off, ws->d_hit_off_cap, size_t(nq) + 1);
          t_sync = Clock::now();
          uint64_t n_hits = locate_chunk_count(r, ws, nq, ws->d_hit_off, st);
          t_search += since(t_sync);
          // results come back through the workspace's pinned buffer (offsets, then hits) and are copied into the
          // caller-bound arrays after the chunk's wait: a device-to-host copy into pageable memory would block this
          // thread for the whole transfer.  The first chunk sizes both arrays for the whole segment.
          const double seg_reads = plan.consumed ? double(out.file_bytes) / (double(plan.consumed) / double(nq)) * 1.02 + double(nq) : 0.0;
          if (out.n_reads == 0 && seg_reads > 0) out.hit_off.reserve(size_t(seg_reads) + 1024);
          out.hit_off.resize(out.n_reads + nq + 1);
          uint64_t* dst_off = out.hit_off.data() + out.n_reads;
          const size_t off_bytes = (size_t(nq) + 1) * 8;
          Workspace::grow_host(ws->h_out, ws->h_out_cap, off_bytes + size_t(n_hits) * 16);
          CU(cudaMemcpyAsync(ws->h_out, ws->d_hit_off, off_bytes, cudaMemcpyDeviceToHost, st));
          if (n_hits) {
            uint64_t* d_hits = locate_chunk_walk(r, ws, nq, n_hits, flags, ws->d_hit_off, st);
            if (out.n_hits + n_hits > out.hits_cap) {
              uint64_t cap = std::max<uint64_t>(out.n_hits + n_hits, out.hits_cap * 2);
              if (out.hits_cap == 0 && seg_reads > 0)  // hits per read of the first chunk, over the whole segment
                cap = std::max<uint64_t>(cap, uint64_t(seg_reads * (double(n_hits) / double(nq)) * 1.05) + 1024);
              populate_done();
              void* np = realloc(out.hits, cap * sizeof(awry_hit));
              if (!np) {
                cudaFreeAsync(d_hits, st);
                fail(AWRY_ERR_NOMEM, "out of host memory for %llu hits", (unsigned long long)cap);
              }
              const bool first_alloc = out.hits_cap == 0;
              out.hits = static_cast<awry_hit*>(np);
              out.hits_cap = cap;
#ifdef MADV_POPULATE_WRITE
              if (first_alloc && cap * sizeof(awry_hit) >= (32u << 20)) {
                const uintptr_t lo = (reinterpret_cast<uintptr_t>(np) + 4095) & ~uintptr_t(4095);
                const uintptr_t hi = (reinterpret_cast<uintptr_t>(np) + cap * sizeof(awry_hit)) & ~uintptr_t(4095);
                if (hi > lo) populate = std::thread([lo, hi] { (void)madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_POPULATE_WRITE); });
              }
#else
              (void)first_alloc;
#endif
            }
            CU(cudaMemcpyAsync(ws->h_out + off_bytes, d_hits, n_hits * 16, cudaMemcpyDeviceToHost, st));
            cudaFreeAsync(d_hits, st);
            g_prof.d2h += n_hits * 16;
          }
          t_sync = Clock::now();
          CU(cudaStreamSynchronize(st));
          t_search += since(t_sync);
          const auto t_copy = Clock::now();
          g_prof.d2h += (nq + 1) * 8;
          if (n_hits) memcpy(out.hits + out.n_hits, ws->h_out + off_bytes, n_hits * 16);
          const uint64_t* src_off = reinterpret_cast<const uint64_t*>(ws->h_out);
          for (uint64_t i = 0; i <= nq; i++) dst_off[i] = src_off[i] + out.n_hits;
          t_copy_out += since(t_copy);
          out.n_hits += n_hits;
        }
        if (*ws->h_flag != ~0ull && bad_read) *bad_read = out.n_reads + (*ws->h_flag >> 1);
        if (*ws->h_flag != ~0ull)
          fail(AWRY_ERR_INVALID_QUERY,
               "read %llu of %s is empty or contains a sentinel ('$'/'#'): the reference panics on it "
               "(fm_index.rs:406, bwt.rs:127)",
               (unsigned long long)(out.n_reads + (*ws->h_flag >> 1)), path);
        out.n_reads += nq;
        out.n_bases += plan.seq_bytes;
      }
      if (eof) break;
      if (!have_next) stage_in(c + 1, true, nxt);
      in = nxt;
    }
  } catch (...) {
    cleanup();
    free(out.hits);
    out.hits = nullptr;
    throw;
  }
  cleanup();
  if (trace)
    fprintf(stderr, "[trace] reads file, replica %zu: %.1f ms in all; main thread waited %.1f ms for the reader, %.1f ms for the "
            "line split, %.1f ms for the record split, %.1f ms for search + results (+ %.1f ms copying hits out); reader "
            "thread read for %.1f ms and waited %.1f ms for a free buffer\n", ri, since(t_call) * 1e3, t_slot * 1e3,
            t_lines * 1e3, t_parse * 1e3, t_search * 1e3, t_copy_out * 1e3, t_pread * 1e3, t_ring * 1e3);
}

// ---- one reads file over SEVERAL replicas (plain files) ----
// The file is cut into one segment per replica at record starts found on the host (a window behind every
// cut is scanned: FASTA -- a line that begins with '>'; FASTQ -- a line that begins with '@' whose next line
// but one begins with '+', which a quality line that happens to begin with '@' never satisfies), every replica
// runs the single-replica pipeline above on its segment with its own reader thread, and the results are
// concatenated in file order.  A gzip stream cannot be entered in the middle and stays on replica 0; so does a
// file too small to be worth splitting, and one whose record starts are not found (multi-line FASTQ).
bool find_record_start(int fd, uint64_t from, uint64_t fsize, bool fastq, uint64_t& at) {
  const uint64_t W = 4u << 20;
  std::vector<char> buf(size_t(std::min<uint64_t>(W, fsize - from)));
  if (buf.empty()) return false;
  ssize_t got = pread(fd, buf.data(), buf.size(), off_t(from));
  if (got <= 0) return false;
  const size_t n = size_t(got);
  auto line_end = [&](size_t p) {  // index of the '\n' ending the line that contains p, or n
    const void* q = memchr(buf.data() + p, '\n', n - p);
    return q ? size_t(static_cast<const char*>(q) - buf.data()) : n;
  };
  size_t p = line_end(0);  // never accept position 0: we do not know what precedes it
  while (p < n) {
    const size_t ls = p + 1;  // a line start
    if (ls >= n) return false;
    if (!fastq) {
      if (buf[ls] == '>') {
        at = from + ls;
        return true;
      }
    } else if (buf[ls] == '@') {
      const size_t e1 = line_end(ls);
      const size_t e2 = e1 < n ? line_end(e1 + 1) : n;
      if (e2 + 1 < n && buf[e2 + 1] == '+') {
        at = from + ls;
        return true;
      }
      if (e2 + 1 >= n) return false;  // window exhausted before the candidate could be confirmed
    }
    p = line_end(ls);
  }
  return false;
}

void run_reads_file_multi(const awry_index* ix, const char* path, bool locate, uint32_t flags, ReadsOut& out) {
  const size_t nr = ix->reps.size();
  const char* off = getenv("AWRY_B200_READS_REPLICAS");  // =1: always replica 0 (A/B runs)
  uint64_t min_seg = 128ull << 20;
  if (const char* e = getenv("AWRY_B200_READS_MIN_SEGMENT")) min_seg = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));
  std::vector<uint64_t> cuts;
  int fastq = -1;
  if (nr > 1 && !(off && off[0] == '1')) {
    int fd = open(path, O_RDONLY);
    if (fd >= 0) {
      struct stat sb;
      unsigned char head[4096];
      ssize_t got = pread(fd, head, sizeof head, 0);
      if (fstat(fd, &sb) == 0 && got >= 2 && !(head[0] == 0x1f && head[1] == 0x8b) && uint64_t(sb.st_size) >= 2 * min_seg) {
        const uint64_t fsize = uint64_t(sb.st_size);
        uint64_t data_start = 0;
        for (ssize_t i = 0; i < got; i++) {
          unsigned char c = head[i];
          if (c == '\n' || c == '\r' || c == ' ' || c == '\t') continue;
          fastq = c == '@' ? 1 : c == '>' ? 0 : -1;
          data_start = uint64_t(i);
          break;
        }
        if (fastq >= 0) {
          const size_t nseg = size_t(std::min<uint64_t>(nr, fsize / min_seg));
          cuts.push_back(data_start);
          bool ok = nseg > 1;
          for (size_t i = 1; i < nseg && ok; i++) {
            uint64_t at = 0;
            ok = find_record_start(fd, data_start + (fsize - data_start) * i / nseg, fsize, fastq == 1, at) && at > cuts.back();
            if (ok) cuts.push_back(at);
          }
          if (ok)
            cuts.push_back(fsize);
          else
            cuts.clear();
        }
      }
      close(fd);
    }
  }
  if (cuts.size() < 3) {  // one segment: the whole file on replica 0 (also: gzip, tiny files, unknown formats)
    run_reads_file(ix, path, locate, flags, out);
    return;
  }
  const size_t nseg = cuts.size() - 1;
  std::vector<ReadsOut> outs(nseg);
  std::vector<int> codes(nseg, 0);
  std::vector<std::string> msgs(nseg);
  std::vector<uint64_t> bad(nseg, ~0ull);
  std::vector<std::thread> th;
  for (size_t i = 0; i < nseg; i++)
    th.emplace_back([&, i] {
      try {
        run_reads_file(ix, path, locate, flags, outs[i], i, cuts[i], cuts[i + 1], fastq, &bad[i]);
      } catch (const ApiError& e) {
        codes[i] = e.code;
        msgs[i] = e.what();
      } catch (const std::exception& e) {
        codes[i] = AWRY_ERR_INVALID_ARG;
        msgs[i] = e.what();
      }
    });
  for (auto& t : th) t.join();
  uint64_t reads_before = 0;
  for (size_t i = 0; i < nseg; i++) {
    if (codes[i]) {
      for (auto& o : outs) free(o.hits);
      if (codes[i] == AWRY_ERR_INVALID_QUERY && bad[i] != ~0ull)  // file-wide read number (the segments before it are complete)
        fail(AWRY_ERR_INVALID_QUERY,
             "read %llu of %s is empty or contains a sentinel ('$'/'#'): the reference panics on it (fm_index.rs:406, bwt.rs:127)",
             (unsigned long long)(reads_before + bad[i]), path);
      fail(codes[i], "%s", msgs[i].c_str());
    }
    reads_before += outs[i].n_reads;
  }
  // concatenate in file order
  uint64_t n_reads = 0, n_hits = 0;
  for (auto& o : outs) {
    n_reads += o.n_reads;
    n_hits += o.n_hits;
    out.n_bases += o.n_bases;
    out.file_bytes += o.file_bytes;
  }
  if (!locate) {
    out.counts.resize(n_reads);
    uint64_t at = 0;
    for (auto& o : outs) {
      if (o.n_reads) memcpy(out.counts.data() + at, o.counts.data(), o.n_reads * 8);
      at += o.n_reads;
    }
  } else {
    out.hit_off.assign(n_reads + 1, 0);
    out.hits = n_hits ? static_cast<awry_hit*>(malloc(n_hits * sizeof(awry_hit))) : nullptr;
    if (n_hits && !out.hits) {
      for (auto& o : outs) free(o.hits);
      fail(AWRY_ERR_NOMEM, "out of host memory for %llu hits", (unsigned long long)n_hits);
    }
    uint64_t at = 0, hbase = 0;
    for (auto& o : outs) {
      for (uint64_t i = 0; i < o.n_reads; i++) out.hit_off[at + i] = hbase + o.hit_off[i];
      if (o.n_hits) memcpy(out.hits + hbase, o.hits, o.n_hits * sizeof(awry_hit));
      free(o.hits);
      o.hits = nullptr;
      at += o.n_reads;
      hbase += o.n_hits;
    }
    out.hit_off[n_reads] = n_hits;
    out.hits_cap = n_hits;
  }
  out.n_reads = n_reads;
  out.n_hits = n_hits;
}

}  // namespace

extern "C" {

int awry_count_reads_file(const awry_index* ix, const char* path, uint64_t** counts, uint64_t* n_reads) {
  return guarded([&] {
    need(ix);
    if (!path || !counts || !n_reads) fail(AWRY_ERR_INVALID_ARG, "null argument");
    *counts = nullptr;
    *n_reads = 0;
    ReadsOut out;
    run_reads_file_multi(ix, path, false, 0, out);
    out.counts.resize(std::max<size_t>(1, size_t(out.n_reads)));  // (never a null array, even for an empty file)
    *counts = out.counts.release();
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
    run_reads_file_multi(ix, path, true, flags, out);
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

}  // extern "C"
