// build_api.cu -- entry points of the GPU index construction (device side and file reader: build.cu):
// FmIndex::new (/root/reference/src/fm_index.rs:142-268) and FmIndex::save (fm_index_file.rs:42-106).
#include "host.hpp"

using namespace awry;
using namespace awry::host;

namespace {

// Sequential writer of an `.awry` v1 file (FmIndex::save, fm_index_file.rs:42-106; sequence index:
// sequence_index.rs:144-152), shared by awry_index_build (arrays fresh from the construction) and
// awry_index_save (arrays re-derived from the device layout).  Write errors are latched and reported by finish().
struct AwryFileOut {
  FILE* f = nullptr;
  bool ok = true;
  std::string path;
  explicit AwryFileOut(const char* p) : path(p) {
    f = fopen(p, "wb");  // create + truncate, as OpenOptions in fm_index_file.rs:43-47
    if (!f) fail(AWRY_ERR_IO, "cannot create %s: %s", p, strerror(errno));
    setvbuf(f, nullptr, _IOFBF, 8u << 20);
  }
  ~AwryFileOut() {
    if (f) fclose(f);
  }
  AwryFileOut(const AwryFileOut&) = delete;
  AwryFileOut& operator=(const AwryFileOut&) = delete;

  void put(const void* p, size_t nbytes) {
    if (ok && nbytes && fwrite(p, 1, nbytes, f) != nbytes) ok = false;
  }
  void put_header(uint64_t version, uint64_t ratio, uint64_t bwt_len, uint64_t alphabet) {
    put("AWRY-Index\n", 11);  // fm_index_file.rs:18,47
    uint64_t hdr[4] = {version, ratio, bwt_len, alphabet};  // generate_file_header, fm_index_file.rs:165-181
    put(hdr, sizeof hdr);
  }
  // n_words u64 that live (or are produced) on `device` -> file through a pinned double buffer: the D2H copy
  // of chunk i+1 overlaps the fwrite of chunk i.  produce(chunk, first_word, n, stream) returns the device
  // address of words [first_word, first_word + n) and may launch the kernel that makes them on `stream`.
  template <class Produce>
  void put_device(int device, uint64_t n_words, uint64_t chunk_words, Produce&& produce) {
    if (n_words == 0) return;
    DeviceGuard dg(device);
    const uint64_t CH = std::min(chunk_words, n_words);
    struct Bufs {
      uint64_t* hb[2] = {nullptr, nullptr};
      cudaStream_t cs = nullptr;
      cudaEvent_t ev[2] = {nullptr, nullptr};
      ~Bufs() {
        if (cs) cudaStreamSynchronize(cs);
        for (int i = 0; i < 2; i++) {
          cudaFreeHost(hb[i]);
          if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (cs) cudaStreamDestroy(cs);
      }
    } b;
    CU(cudaStreamCreateWithFlags(&b.cs, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CU(cudaHostAlloc(reinterpret_cast<void**>(&b.hb[i]), CH * 8, cudaHostAllocDefault));
      CU(cudaEventCreateWithFlags(&b.ev[i], cudaEventDisableTiming));
    }
    const uint64_t n_chunks = (n_words + CH - 1) / CH;
    auto issue = [&](uint64_t c) {
      const uint64_t w0 = c * CH, nw = std::min(CH, n_words - w0);
      const uint64_t* d_src = produce(c, w0, nw, b.cs);
      CU(cudaMemcpyAsync(b.hb[c & 1], d_src, nw * 8, cudaMemcpyDeviceToHost, b.cs));
      CU(cudaEventRecord(b.ev[c & 1], b.cs));
    };
    issue(0);
    for (uint64_t c = 0; c < n_chunks; c++) {
      CU(cudaEventSynchronize(b.ev[c & 1]));
      if (c + 1 < n_chunks) issue(c + 1);
      put(b.hb[c & 1], std::min(CH, n_words - c * CH) * 8);
    }
  }
  // 1 byte k, then (card-2)^k entries populated the way kmer_lookup_table.rs:121-167 does;
  // computed on replica 0 from the rank blocks, so a handle needs no copy of a table nobody reads
  void put_table_section(const awry_index* ix, uint32_t k) {
    uint8_t kb = uint8_t(k);
    put(&kb, 1);
    Replica& r = *ix->reps[0];
    DeviceGuard dg(r.device);
    const uint64_t n_entries = ipow(uint64_t(ix->card - 2), k), CH = 1u << 22;
    ulonglong2* d_buf = nullptr;
    CU(cudaMalloc(reinterpret_cast<void**>(&d_buf), std::min(n_entries, CH) * 16));
    std::unique_ptr<ulonglong2, void (*)(ulonglong2*)> d_guard(d_buf, [](ulonglong2* p) { cudaFree(p); });
    std::vector<uint64_t> h_buf(2 * std::min(n_entries, CH));
    IndexView v = r.view;
    v.kmer_len = 0;
    for (uint64_t first = 0; first < n_entries && ok; first += CH) {
      uint64_t cnt = std::min(CH, n_entries - first);
      cudaError_t e = launch_ref_table(v, first, cnt, k, d_buf, nullptr);
      if (e == cudaSuccess) e = cudaMemcpy(h_buf.data(), d_buf, cnt * 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) fail(AWRY_ERR_CUDA, "k-mer table kernel failed: %s", cudaGetErrorString(e));
      put(h_buf.data(), cnt * 16);
    }
  }
  void put_sequence_index(const std::vector<uint64_t>& starts, const std::vector<std::string>& headers, uint64_t n_seqs) {
    put(&n_seqs, 8);  // sequence_index.rs:144-152
    for (uint64_t i = 0; i < n_seqs; i++) {
      const uint64_t st = i < starts.size() ? starts[i] : 0, hl = i < headers.size() ? headers[i].size() : 0;
      put(&st, 8);
      put(&hl, 8);
      if (hl) put(headers[i].data(), hl);
    }
  }
  void finish() {
    if (fflush(f) != 0) ok = false;
    if (fclose(f) != 0) ok = false;
    f = nullptr;
    if (!ok) fail(AWRY_ERR_IO, "write to %s failed", path.c_str());
  }
};

}  // namespace

extern "C" {

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
    if (n + 1 >= (1ull << 56)) fail(AWRY_ERR_UNSUPPORTED, "text of %llu symbols", (unsigned long long)n);
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
    if (bwt_len >= (1ull << 56)) fail(AWRY_ERR_UNSUPPORTED, "text of %llu symbols", (unsigned long long)n);
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
      AwryFileOut fo(a->output_file_src);
      fo.put_header(1, ratio, bwt_len, uint64_t(alphabet));
      fo.put_device(dp.device, dp.n_block_words, (64u << 20) / 8,
                    [&](uint64_t, uint64_t w0, uint64_t, cudaStream_t) { return dp.d_blocks + w0; });
      fo.put(prefix.data(), prefix.size() * 8);
      fo.put_device(dp.device, sa_word_len(bwt_len, ratio), (64u << 20) / 8,
                    [&](uint64_t, uint64_t w0, uint64_t, cudaStream_t) { return dp.d_sa_words + w0; });
      fo.put_table_section(ix, k);
      fo.put_sequence_index(starts, headers, starts.size());
      fo.finish();
      tick("write .awry file");
    }
    if (out) *out = holder.release();
  });
}

int awry_index_save(const awry_index* ix, const char* path) {
  return guarded([&] {
    need(ix);
    if (!path) fail(AWRY_ERR_INVALID_ARG, "null argument");
    Replica& r = *ix->reps[0];
    const uint64_t n_rb = (ix->bwt_len + 255) / 256, wpb = ix->alphabet == AWRY_NUCLEOTIDE ? 20 : 44;
    AwryFileOut fo(path);
    fo.put_header(ix->version, ix->sa_ratio, ix->bwt_len, uint64_t(ix->alphabet));
    {
      // the handle keeps only the device layout: the bwt.rs blocks are re-derived chunk by chunk
      DeviceGuard dg(r.device);
      const uint64_t chunk_blocks = std::min<uint64_t>(1u << 18, n_rb);
      struct Tmp {
        uint64_t* d[2] = {nullptr, nullptr};
        ~Tmp() {
          cudaFree(d[0]);
          cudaFree(d[1]);
        }
      } tmp;
      for (auto& d : tmp.d) CU(cudaMalloc(reinterpret_cast<void**>(&d), chunk_blocks * wpb * 8));
      fo.put_device(r.device, n_rb * wpb, chunk_blocks * wpb, [&](uint64_t c, uint64_t w0, uint64_t nw, cudaStream_t cs) {
        CU(launch_untranspose(ix->alphabet, r.d_blocks, r.dollar_row, w0 / wpb, nw / wpb, tmp.d[c & 1], r.d_sb, ix->sb_shift, cs));
        return static_cast<const uint64_t*>(tmp.d[c & 1]);
      });
    }
    fo.put(ix->prefix_sums, size_t(ix->card + 1) * 8);
    fo.put_device(r.device, ix->n_sa_words, (64u << 20) / 8,
                  [&](uint64_t, uint64_t w0, uint64_t, cudaStream_t) { return static_cast<const uint64_t*>(r.d_sa + w0); });
    fo.put_table_section(ix, ix->kmer_len_file);
    fo.put_sequence_index(ix->seq_starts, ix->headers, ix->headers.size());
    fo.finish();
  });
}

int awry_build_index_file(const awry_build_args* a) {
  if (a && !a->output_file_src) {
    return guarded([&] { fail(AWRY_ERR_INVALID_ARG, "output_file_src is null"); });
  }
  return awry_index_build(a, nullptr, 0, nullptr);
}

}  // extern "C"
