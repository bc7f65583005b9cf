// api.cu -- index lifetime of libawry_b200: the `.awry` v1 loader, device replicas, getters, the two
// single-step functions and the instrumentation entry points of include/awry_b200.h.
// (batch.cu: count / locate pipelines; reads_api.cu: FASTQ front-end; build_api.cu: GPU construction.)
//
// Reference host code this stands in for (paths under /root/reference/src):
//   FmIndex::load / read_fm_index_by_version_number   fm_index_file.rs:132-160, :184-287
//   KmerLookupTable::from_file (skipped, Q1/Q2)        kmer_lookup_table.rs:55-77
//   SequenceIndex::from_file                           sequence_index.rs:155-183
#include "host.hpp"

using namespace awry;
using namespace awry::host;

namespace awry {
namespace host {

thread_local char g_err[1024] = "";
ProfState g_prof;
std::atomic<uint64_t> g_variant_bits{uint64_t(1)};  // lanes 0, tpb 0, blocks_per_sm 0, slots -1 (see host.hpp)
std::atomic<int> g_host_pack{-1};
std::atomic<int> g_locate_variant{0};
std::atomic<int> g_count_variant{0};

[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[900];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw ApiError(code, buf);
}

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

// threads of a parallel positional read (AWRY_B200_IO_THREADS, default min(8, cores))
unsigned io_threads() {
  static const unsigned v = [] {
    if (const char* e = getenv("AWRY_B200_IO_THREADS")) return unsigned(std::min(64l, std::max(1l, strtol(e, nullptr, 10))));
    return std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  }();
  return v;
}


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



// sizes of the arrays, in elements (AWRY_CHK in the checked build; the allocations carry 256 B of slack on top);
// called again whenever one of the derived arrays has been built
void set_limits(awry_index* ix, Replica& r) {
  IndexView& v = r.view;
  v.n_blocks_u4 = r.bytes_blocks / 16;
  v.n_pair_u4 = r.d_pair ? pair_block_count(ix->bwt_len) * PAIR_BLOCK_UINT4 : 0;
  v.n_table = ix->wide ? 0 : (r.d_table ? table_entries(ix->alphabet, ix->kmer_len_dev) : 0);
  v.n_sa_words = ix->n_sa_words + 2;
  v.n_full_sa = r.d_full_sa ? ix->bwt_len : 0;
  v.n_rtext = r.d_rtext ? rtext_bytes(ix->alphabet, ix->bwt_len) : 0;
  v.n_walk_u4 = r.d_walk ? walk_block_count(ix->bwt_len) * WALK_BLOCK_UINT4 : 0;
  v.n_walk_rank = r.d_walk ? walk_block_count(ix->bwt_len) + 1 : 0;
  v.n_pos_samples = r.d_walk && ix->lean_ratio ? (ix->bwt_len + ix->lean_ratio - 1) / ix->lean_ratio : 0;
  WideView& w = r.wview;
  w.n_blocks_u4 = r.bytes_blocks / 16;
  w.n_table = r.d_table_w ? table_entries(ix->alphabet, ix->kmer_len_dev) : 0;
  w.n_sa_words = ix->n_sa_words + 2;
  w.n_sb = ix->n_superblocks();
}

void set_view_constants(awry_index* ix, Replica& r) {
  IndexView& v = r.view;
  v.blocks = r.d_blocks;
  v.sa_words = r.d_sa;
  v.table = r.d_table;
  v.seq_starts = r.d_seq_starts;
  v.pair_blocks = r.d_pair;
  v.full_sa = r.d_full_sa;
  v.rtext = r.d_rtext;
  v.walk_blocks = r.d_walk;
  v.walk_rank = r.d_walk_rank;
  v.pos_samples = r.d_pos_samples;
  v.lean_ratio = ix->lean_ratio;
  set_limits(ix, r);
  for (int i = 0; i < 16; i++) v.c2[i] = r.c2[i];
  v.bwt_len = ix->wide ? 0xffffffffu : uint32_t(ix->bwt_len);
  v.dollar_row = ix->wide ? 0xffffffffu : uint32_t(r.dollar_row);
  v.wide = ix->wide ? &r.wview : nullptr;
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
  if (ix->wide) {
    WideView& w = r.wview;
    w.blocks = r.d_blocks;
    w.sa_words = r.d_sa;
    w.table = r.d_table_w;
    w.seq_starts = r.d_seq_starts;
    w.sb_counts = r.d_sb;
    w.bwt_len = ix->bwt_len;
    w.dollar_row = r.dollar_row;
    w.sa_ratio = v.sa_ratio;
    w.sa_pow2 = v.sa_pow2;
    w.sa_ratio_shift = v.sa_ratio_shift;
    w.sa_bits = v.sa_bits;
    w.kmer_len = v.kmer_len;
    w.n_seqs = v.n_seqs;
    w.alphabet = v.alphabet;
    w.sb_shift = ix->sb_shift;
    set_limits(ix, r);
    for (int i = 0; i < 24; i++) w.c_lo[i] = 1, w.c_hi[i] = 0;
    if (ix->alphabet == AWRY_NUCLEOTIDE) {
      static const int ref_of_dsym[6] = {1, 2, 3, 5, 4, 0};  // A C G T N $
      for (int d = 0; d < 6; d++) {
        w.c_lo[d] = ix->prefix_sums[ref_of_dsym[d]];
        w.c_hi[d] = ix->prefix_sums[ref_of_dsym[d] + 1] - 1;
      }
    } else {
      for (int s2 = 0; s2 < 22; s2++) {
        w.c_lo[s2] = ix->prefix_sums[s2];
        w.c_hi[s2] = ix->prefix_sums[s2 + 1] - 1;
      }
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
  unsigned long long* d_dollar = nullptr;
  CU(cudaMalloc(reinterpret_cast<void**>(&d_dollar), 8));
  CU(cudaMemset(d_dollar, 0xff, 8));
  // wide index: the u32 block counts become relative to superblocks of 2^sb_shift rows; the absolute counts of
  // a superblock are the milestones of its first reference block (fm_index.rs:212-217), picked up as the
  // blocks stream by
  const uint32_t sbs = ix->sb_shift;
  const uint64_t rb_per_sb = 1ull << (sbs - 8);
  if (ix->wide) {
    CU(cudaMalloc(reinterpret_cast<void**>(&r.d_sb), ix->n_superblocks() * SB_STRIDE * 8));
    CU(cudaMemset(r.d_sb, 0, ix->n_superblocks() * SB_STRIDE * 8));
  }

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
  if (src_blocks_then_rest.dev_blocks) {
    if (ix->wide) CU(launch_gather_superblocks(ix->alphabet, src_blocks_then_rest.dev_blocks, 0, n_ref_blocks, sbs, r.d_sb, st));
    CU(launch_transpose(ix->alphabet, src_blocks_then_rest.dev_blocks, 0, n_ref_blocks, ix->bwt_len, r.d_blocks, d_dollar,
                        r.d_sb, sbs, st));
  }
  for (uint64_t b0 = 0; b0 < n_ref_blocks && !src_blocks_then_rest.dev_blocks; b0 += chunk_blocks, slot ^= 1) {
    uint64_t nb = std::min(chunk_blocks, n_ref_blocks - b0);
    CU(cudaEventSynchronize(ev[slot]));
    src_blocks_then_rest.read(h_stage[slot], nb * ref_block_bytes, "bwt blocks");
    CU(cudaMemcpyAsync(d_stage[slot], h_stage[slot], nb * ref_block_bytes, cudaMemcpyHostToDevice, st));
    if (ix->wide && ((b0 + rb_per_sb - 1) / rb_per_sb) * rb_per_sb < b0 + nb)  // a superblock starts in this chunk
      CU(launch_gather_superblocks(ix->alphabet, d_stage[slot], b0, nb, sbs, r.d_sb, st));
    CU(launch_transpose(ix->alphabet, d_stage[slot], b0, nb, ix->bwt_len, r.d_blocks, d_dollar, r.d_sb, sbs, st));
    CU(cudaEventRecord(ev[slot], st));
  }
  CU(cudaStreamSynchronize(st));
  unsigned long long dollar = 0;
  CU(cudaMemcpy(&dollar, d_dollar, 8, cudaMemcpyDeviceToHost));
  cudaFree(d_dollar);
  if (dollar == ~0ull) fail(AWRY_ERR_FORMAT, "no sentinel row found in the BWT blocks");

  // prefix sums (fm_index_file.rs:265-270): C[0] = 0, C[1] = 1 (one '$' row), non-decreasing, C[card] = bwt_len --
  // every c_hi = C[c+1] - 1 and every row pointer the kernels form rests on this
  src_blocks_then_rest.read(ix->prefix_sums, size_t(ix->card + 1) * 8, "prefix sums");
  if (ix->prefix_sums[ix->card] != ix->bwt_len)
    fail(AWRY_ERR_FORMAT, "prefix sums do not add up to bwt_len (%llu vs %llu)",
         (unsigned long long)ix->prefix_sums[ix->card], (unsigned long long)ix->bwt_len);
  if (ix->prefix_sums[0] != 0 || ix->prefix_sums[1] != 1)
    fail(AWRY_ERR_FORMAT, "prefix sums do not start with 0, 1 (exactly one sentinel row)");
  for (int c = 0; c < ix->card; c++)
    if (ix->prefix_sums[c + 1] < ix->prefix_sums[c]) fail(AWRY_ERR_FORMAT, "prefix sums decrease at symbol %d", c);

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
    // (card-2)^k x 16 bytes must lie inside the file: an unbounded k would overflow the size and the seek
    // would land anywhere
    const uint32_t k_max = ix->alphabet == AWRY_NUCLEOTIDE ? 27 : 12;  // 4^27, 20^12 entries x 16 B < 2^63
    if (k > k_max) fail(AWRY_ERR_FORMAT, "k-mer length %u in the file is implausible", unsigned(k));
    uint64_t n_entries = ipow(uint64_t(ix->card - 2), k);
    if (src_blocks_then_rest.f) {
      struct stat sb;
      const off_t here = ftello(src_blocks_then_rest.f);
      if (fstat(fileno(src_blocks_then_rest.f), &sb) != 0 || here < 0 ||
          uint64_t(sb.st_size) < uint64_t(here) + n_entries * 16 + 8)
        fail(AWRY_ERR_FORMAT, "file too short for its k-mer table section (k = %u)", unsigned(k));
      if (fseeko(src_blocks_then_rest.f, off_t(n_entries * 16), SEEK_CUR) != 0)
        fail(AWRY_ERR_IO, "cannot seek past the kmer table");
    }
  }

  // sequence starts on the device (filled by the caller into ix->seq_starts before this returns
  // for parts; for files they follow the table -- read by the caller after this function)
  r.dollar_row = dollar;
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
thread_local bool g_skip_accelerators = false;  // declared in host.hpp

void finish_replica0(awry_index* ix, Replica& r) {
  DeviceGuard dg(r.device);
  if (ix->seq_starts.empty()) ix->seq_starts.push_back(0);
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_seq_starts), ix->seq_starts.size() * 8));
  CU(cudaMemcpy(r.d_seq_starts, ix->seq_starts.data(), ix->seq_starts.size() * 8, cudaMemcpyHostToDevice));
  CU(cudaMalloc(reinterpret_cast<void**>(&r.d_async_flag), 8));
  CU(cudaMemset(r.d_async_flag, 0xff, 8));
  keep_pool_memory(r.device);

  // device seed table: the table only accelerates, results are those of the plain backward search either way, so
  // its depth is the library's choice: the file's k (capped at 14), or deeper when the index is large enough to
  // make use of it -- nucleotide: the deepest k <= 15 with 4^k <= rows / 2 whose table takes at most an eighth of
  // the free memory (k = 15, 8.6 GB, at 3.1 G rows: a seeded interval is then ~3 rows wide and a wave of the count
  // kernel needs one or two pair steps instead of three: 1.70 / 1.48 / 1.34 / 1.19 ms per 10 M x 150-bp reads
  // with k = 13 / 14 / 15 / 16, profiles/r02_s11_seed_depth.log; k = 16 would be 34 GB).  AWRY_B200_KMER_DEV
  // sets it (up to 16).
  const bool dna = ix->alphabet == AWRY_NUCLEOTIDE;
  uint32_t cap = dna ? 14 : 6;
  uint32_t k = std::min<uint32_t>(ix->kmer_len_file, cap);
  const uint32_t k_file = k;
  if (!ix->wide && k > 0) {  // (protein: up to k = 6, 0.5 GB -- one block read less per peptide)
    uint32_t k_auto = 0;
    while (k_auto < (dna ? 15u : cap) && table_entries(ix->alphabet, k_auto + 1) <= ix->bwt_len / 2) k_auto++;
    k = std::max(k, k_auto);
  }
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  if (const char* e = getenv("AWRY_B200_KMER_DEV")) k = std::min<uint32_t>(uint32_t(std::max(0l, strtol(e, nullptr, 10))), dna ? 16u : cap);  // experiments
  const size_t entry_bytes = ix->wide ? 16 : 8;
  if (!getenv("AWRY_B200_KMER_DEV"))
    while (k > k_file && table_entries(ix->alphabet, k) * entry_bytes > free_b / 8) k--;  // the deeper table is a luxury
  while (k > 0 && table_entries(ix->alphabet, k) * entry_bytes > free_b / 4) k--;
  ix->kmer_len_dev = k;
  set_view_constants(ix, r);
  if (k > 0 && ix->wide) {
    r.bytes_table = table_entries(ix->alphabet, k) * entry_bytes;
    CU(cudaMalloc(reinterpret_cast<void**>(&r.d_table_w), r.bytes_table));
    r.wview.table = r.d_table_w;
    set_limits(ix, r);
    WideView v = r.wview;
    v.kmer_len = 0;
    CU(launch_build_table_wide(v, r.d_table_w, k, nullptr));
    CU(cudaDeviceSynchronize());
  } else if (k > 0) {
    r.bytes_table = table_entries(ix->alphabet, k) * 8;
    CU(cudaMalloc(reinterpret_cast<void**>(&r.d_table), r.bytes_table));
    r.view.table = r.d_table;
    set_limits(ix, r);
    IndexView v = r.view;
    v.kmer_len = 0;
    CU(launch_build_table(v, r.d_table, k, nullptr));
    CU(cudaDeviceSynchronize());
  }
  if (ix->wide) return;  // the pair index, the unsampled array and the walk blocks are 32-bit structures
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
      set_limits(ix, r);
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
      set_limits(ix, r);
    }
  }
  // Nucleotide: the text itself, reversed, 4 bits per symbol (layout.cuh), read back out of BWT + unsampled array.
  // With it the count kernel finishes a query whose interval has narrowed to one row with one suffix-array read
  // and one or two lines of text instead of a block read per two symbols.  n / 2 bytes; AWRY_B200_TEXT=0 never.
  const char* tx = getenv("AWRY_B200_TEXT");
  // Protein: the same with one byte per symbol (bwt_len bytes).
  if (r.d_full_sa && (ix->alphabet != AWRY_NUCLEOTIDE || r.d_pair) && !(tx && tx[0] == '0') && !g_skip_accelerators) {
    const size_t bytes = rtext_bytes(ix->alphabet, ix->bwt_len);
    CU(cudaMemGetInfo(&free_b, &total_b));
    if (bytes < free_b / 3) {
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_rtext), bytes));
      CU(build_rtext(r.view, r.d_rtext, nullptr));
      CU(cudaDeviceSynchronize());
      r.bytes_rtext = bytes;
      r.view.rtext = r.d_rtext;
      set_limits(ix, r);
    }
  }
  // Memory-lean bounded locate (nucleotide): walk blocks + position-sampled suffix array, 4.57 + 32/ratio bits
  // per row.  Built when the unsampled array is not (it did not fit, or AWRY_B200_FULL_SA=0), or on request
  // (AWRY_B200_LEAN_SA=1: both, for A/B runs; =0: never -- locate then LF-walks to the file's row samples).
  const char* ls = getenv("AWRY_B200_LEAN_SA");
  const bool lean_never = ls && ls[0] == '0', lean_force = ls && ls[0] == '1';
  if (ix->alphabet == AWRY_NUCLEOTIDE && !g_skip_accelerators && ix->sa_ratio > 1 && !lean_never &&
      (lean_force || r.d_full_sa == nullptr)) {
    // The sampling distance is ours to choose (nothing of this is in the file): 4 unless the file samples more
    // densely -- mean walk 1.5 steps, 12.6 bits per row -- doubled while the arrays would take more than a third
    // of the free memory.  AWRY_B200_LEAN_RATIO overrides.
    const uint64_t nb = walk_block_count(ix->bwt_len);
    CU(cudaMemGetInfo(&free_b, &total_b));
    uint64_t lr = std::min<uint64_t>(ix->sa_ratio, 4);
    if (const char* e = getenv("AWRY_B200_LEAN_RATIO")) lr = std::min<uint64_t>(1u << 20, std::max<uint64_t>(2, strtoull(e, nullptr, 10)));
    auto lean_bytes = [&](uint64_t ratio) {
      return size_t(nb) * 128 + size_t(nb + 1) * 4 + size_t((ix->bwt_len + ratio - 1) / ratio + 4) * 4;
    };
    while (lr < (1u << 20) && lean_bytes(lr) > free_b / 3) lr *= 2;
    ix->lean_ratio = uint32_t(lr);
    r.view.lean_ratio = ix->lean_ratio;
    const uint64_t n_pos = (ix->bwt_len + lr - 1) / lr;
    const size_t bytes = lean_bytes(lr);
    if (bytes + (1u << 28) < free_b) {
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_walk), size_t(nb) * 128 + 256));
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_walk_rank), size_t(nb + 1) * 4 + 256));
      CU(cudaMalloc(reinterpret_cast<void**>(&r.d_pos_samples), size_t(n_pos + 4) * 4 + 256));
      set_limits(ix, r);  // (the build kernels themselves run under the checks of the checked build)
      CU(build_lean_sa(r.view, r.d_walk, r.d_walk_rank, r.d_pos_samples, r.sm_count, nullptr));
      r.bytes_lean = bytes;
      r.view.walk_blocks = r.d_walk;
      set_limits(ix, r);
      r.view.walk_rank = r.d_walk_rank;
      r.view.pos_samples = r.d_pos_samples;
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
  if (src.bytes_table && ix->wide) {
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_table_w), src.bytes_table));
    CU(cudaMemcpyPeer(dst.d_table_w, dst.device, src.d_table_w, src.device, src.bytes_table));
  } else if (src.bytes_table) {
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_table), src.bytes_table));
    CU(cudaMemcpyPeer(dst.d_table, dst.device, src.d_table, src.device, src.bytes_table));
  }
  if (ix->wide) {
    const size_t sbb = ix->n_superblocks() * SB_STRIDE * 8;
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_sb), sbb));
    CU(cudaMemcpyPeer(dst.d_sb, dst.device, src.d_sb, src.device, sbb));
  }
  dst.dollar_row = src.dollar_row;
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
  if (src.bytes_rtext) {
    dst.bytes_rtext = src.bytes_rtext;
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_rtext), src.bytes_rtext));
    CU(cudaMemcpyPeer(dst.d_rtext, dst.device, src.d_rtext, src.device, src.bytes_rtext));
  }
  if (src.bytes_lean) {
    const uint64_t nb = walk_block_count(ix->bwt_len);
    const uint64_t n_pos = (ix->bwt_len + ix->lean_ratio - 1) / ix->lean_ratio;
    dst.bytes_lean = src.bytes_lean;
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_walk), size_t(nb) * 128 + 256));
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_walk_rank), size_t(nb + 1) * 4 + 256));
    CU(cudaMalloc(reinterpret_cast<void**>(&dst.d_pos_samples), size_t(n_pos + 4) * 4 + 256));
    CU(cudaMemcpyPeer(dst.d_walk, dst.device, src.d_walk, src.device, size_t(nb) * 128));
    CU(cudaMemcpyPeer(dst.d_walk_rank, dst.device, src.d_walk_rank, src.device, size_t(nb + 1) * 4));
    CU(cudaMemcpyPeer(dst.d_pos_samples, dst.device, src.d_pos_samples, src.device, size_t(n_pos + 4) * 4));
  }
  set_view_constants(ix, dst);
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
  if (ix->bwt_len >= (1ull << 56)) fail(AWRY_ERR_UNSUPPORTED, "bwt_len %llu >= 2^56", (unsigned long long)ix->bwt_len);
  // 32-bit row pointers and the cooperative kernels while they suffice (every BASELINE config); 64-bit rows
  // (SearchPtr = u64, search.rs:7) past that.  AWRY_B200_WIDE=1 forces the wide path on any index and
  // AWRY_B200_SB_SHIFT shrinks its superblocks, so that small test indexes cross superblock borders.
  ix->wide = ix->bwt_len >= (1ull << 32) - 256;
  if (const char* e = getenv("AWRY_B200_WIDE")) ix->wide = ix->wide || e[0] == '1';
  ix->sb_shift = SB_SHIFT;
  if (const char* e = getenv("AWRY_B200_SB_SHIFT")) ix->sb_shift = uint32_t(std::min(31l, std::max(8l, strtol(e, nullptr, 10))));
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
    ix->balance.push_back(std::make_unique<PackBalance>());
  }
  Replica& r0 = *ix->reps[0];
  build_replica0(ix, r0, src, from_file);
  if (from_file) {
    // sequence index (sequence_index.rs:155-183)
    uint64_t n = 0;
    src.read(&n, 8, "sequence count");
    if (n >= (1ull << 32)) fail(AWRY_ERR_FORMAT, "implausible sequence count");
    for (uint64_t i = 0; i < n; i++) {
      uint64_t start = 0, hl = 0;
      src.read(&start, 8, "sequence start");
      src.read(&hl, 8, "header length");
      if (hl > (1ull << 30)) fail(AWRY_ERR_FORMAT, "implausible header length");
      // the device maps a position to its record by binary search over the starts (sequence_index.rs:108-141)
      if (start >= ix->bwt_len || (i > 0 && start <= ix->seq_starts.back()))
        fail(AWRY_ERR_FORMAT, "sequence start %llu of record %llu is out of order or beyond the text",
             (unsigned long long)start, (unsigned long long)i);
      std::string h(hl, '\0');
      src.read(h.data(), hl, "header");
      ix->seq_starts.push_back(start);
      ix->headers.push_back(std::move(h));
    }
  }
  finish_replica0(ix, r0);
  for (size_t i = 1; i < ix->reps.size(); i++) clone_replica(ix, r0, *ix->reps[i]);
}

}  // namespace host
}  // namespace awry

namespace {

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


}  // namespace

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
    {  // the sections the header implies must fit the file before anything is read or allocated from them
      struct stat sb;
      const uint64_t need_bytes = 43 + ((ix->bwt_len + 255) / 256) * (ix->alphabet == AWRY_NUCLEOTIDE ? 160 : 352) +
                                  uint64_t(ix->card + 1) * 8 + ix->n_sa_words * 8 + 1 + 8;
      if (fstat(fileno(f), &sb) != 0 || uint64_t(sb.st_size) < need_bytes)
        fail(AWRY_ERR_FORMAT, "file is shorter (%lld bytes) than its header implies (>= %llu bytes)",
             (long long)sb.st_size, (unsigned long long)need_bytes);
      // ... and the k-mer table section, whose size hangs on one byte (kmer_lookup_table.rs:55-77)
      uint8_t k = 0;
      if (pread(fileno(f), &k, 1, off_t(need_bytes - 9)) != 1) fail(AWRY_ERR_IO, "cannot read the k-mer length");
      if (k > (ix->alphabet == AWRY_NUCLEOTIDE ? 27 : 12)) fail(AWRY_ERR_FORMAT, "k-mer length %u in the file is implausible", unsigned(k));
      if (uint64_t(sb.st_size) < need_bytes + ipow(uint64_t(ix->card - 2), k) * 16)
        fail(AWRY_ERR_FORMAT, "file is shorter (%lld bytes) than its k-mer table section (k = %u) needs",
             (long long)sb.st_size, unsigned(k));
    }
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
    cudaFree(r.d_table_w);
    cudaFree(r.d_sb);
    cudaFree(r.d_pair);
    cudaFree(r.d_full_sa);
    cudaFree(r.d_rtext);
    cudaFree(r.d_walk);
    cudaFree(r.d_walk_rank);
    cudaFree(r.d_pos_samples);
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
    info->device_bytes_lean_sa = ix->reps[0]->bytes_lean;
    for (size_t i = 0; i < ix->reps.size() && i < 16; i++) info->devices[i] = ix->reps[i]->device;
    info->row_pointer_bits = ix->wide ? 64 : 32;
    info->lean_sa_ratio = ix->reps[0]->bytes_lean ? ix->lean_ratio : 0;
    info->device_bytes_text = ix->reps[0]->bytes_rtext;
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

int awry_initial_range(const awry_index* ix, uint8_t ascii_symbol, awry_range* out) {
  return guarded([&] {
    need(ix);
    if (!out) fail(AWRY_ERR_INVALID_ARG, "null out");
    bool sent = false;
    uint32_t d = ascii_to_dsym(ix, ascii_symbol, &sent);
    // SearchRange::new accepts the sentinel too: [C[0], C[1]-1] (search.rs:43-48)
    if (sent) {
      out->start_ptr = ix->prefix_sums[0];
      out->end_ptr = ix->prefix_sums[1] - 1;
    } else if (ix->wide) {
      out->start_ptr = ix->reps[0]->wview.c_lo[d];
      out->end_ptr = ix->reps[0]->wview.c_hi[d];
    } else {
      out->start_ptr = ix->reps[0]->view.c_lo[d];
      out->end_ptr = ix->reps[0]->view.c_hi[d];
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
      CU(launch_single_update(r.view, range.start_ptr, range.end_ptr, d, reinterpret_cast<uint64_t*>(ws->d_out), ws->st));
      uint64_t res[2];
      CU(cudaMemcpyAsync(res, ws->d_out, 16, cudaMemcpyDeviceToHost, ws->st));
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
      CU(launch_single_backstep(r.view, row, reinterpret_cast<uint64_t*>(ws->d_out), ws->st));
      uint64_t res = 0;
      CU(cudaMemcpyAsync(&res, ws->d_out, 8, cudaMemcpyDeviceToHost, ws->st));
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

}  // extern "C"
