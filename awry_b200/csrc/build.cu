// build.cu -- GPU index construction ("next" after the search path).
//
// What the reference does on the CPU in FmIndex::new (/root/reference/src/fm_index.rs:142-268) on top
// of libsufr: text -> suffix array -> one pass producing BWT blocks (256 rows: bit-planes then u64
// milestones, bwt.rs:12-25), prefix sums and the bit-packed row-sampled suffix array
// (compressed_suffix_array.rs:51-64); and FmIndex::save (fm_index_file.rs:42-106).
// Here: one CUB radix sort of (64-bit packed prefix, position) pairs, ties refined by prefix
// doubling on the device (or on the host when there are only a few), BWT planes by warp ballots,
// milestones by a scan.  Output is the REFERENCE layout / `.awry` v1 file, so the result loads in
// the reference as well as through awry_index_load.  3.1 Gbp in ~1.5-3 s on a B200.
//
// Text model handed to the reference by libsufr (fm_index.rs:148-153, :220-229; libsufr itself is
// not vendored, so this is its documented behaviour): records upper-cased, joined by 'N' / 'X',
// one trailing '$'; suffixes in the order of the symbol indices of alphabet.rs:169-248
// ('$' < A < C < G < N < T), which is byte order for upper-case ACGTN / amino-acid text.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>
#include <thread>
#include <unistd.h>
#include <string>
#include <vector>

#include "build.hpp"

namespace {

#define CU(x)                                                                     \
  do {                                                                            \
    cudaError_t e__ = (x);                                                        \
    if (e__ != cudaSuccess) {                                                     \
      char b__[400];                                                              \
      snprintf(b__, sizeof b__, "%s: %s (line %d)", #x, cudaGetErrorString(e__), __LINE__); \
      throw std::string(b__);                                                     \
    }                                                                             \
  } while (0)

__host__ __device__ inline uint8_t ascii_to_index(int alphabet, uint8_t ch) {
  if (ch >= 'a' && ch <= 'z') ch = uint8_t(ch - 'a' + 'A');
  if (ch == '$' || ch == '#') return 0;
  if (alphabet == 0) return ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : (ch == 'T' || ch == 'U') ? 5 : 4;
  const char* A = "$ACDEFGHIKLMNPQRSTVWXY";
  for (int i = 1; i < 22; i++)
    if (i != 20 && A[i] == char(ch)) return uint8_t(i);
  return 20;
}
__host__ __device__ inline uint8_t index_to_code(int alphabet, uint8_t idx) {  // alphabet.rs:250-330
  const uint8_t D[6] = {0x4, 0x6, 0x5, 0x3, 0x2, 0x1};
  const uint8_t A[22] = {0x00, 0x0c, 0x17, 0x03, 0x06, 0x1e, 0x1a, 0x1b, 0x19, 0x15, 0x1c,
                         0x1d, 0x08, 0x09, 0x04, 0x13, 0x0a, 0x05, 0x16, 0x01, 0x1f, 0x02};
  return alphabet == 0 ? D[idx] : A[idx];
}

// in place ASCII -> symbol index; sym[n] = 0 ('$').  flags[0]: a sentinel character occurs in the
// text; flags[1]: a nucleotide text holds something other than A/C/G/T (then 2-bit keys are out).
__global__ void ascii_to_sym_kernel(int alphabet, uint64_t n, uint8_t* sym, unsigned int* flags) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  bool sent = false, other = false;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i <= n; i += stride) {
    uint8_t x = i == n ? 0 : ascii_to_index(alphabet, sym[i]);
    if (i < n) {
      sent |= x == 0;
      other |= x == 4;
    }
    sym[i] = x;
  }
  if (sent) flags[0] = 1;
  if (other) flags[1] = 1;
}

// key = first SPK symbols of suffix i, BITS bits each, most significant first; positions past
// the '$' read as 0.  BITS == 2 is only used for pure-ACGT text (A0 C1 G2 T3; '$' reads as 0,
// which makes ties that the host fix-up resolves).
template <int BITS>
__global__ void make_keys_kernel(const uint8_t* __restrict__ sym, uint64_t n1, uint64_t first,
                                 uint64_t count, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  constexpr int SPK = 64 / BITS;
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < count; t += stride) {
    uint64_t i = first + t;
    uint64_t key = 0;
#pragma unroll 8
    for (int j = 0; j < SPK; j++) {
      uint32_t s = i + j < n1 ? sym[i + j] : 0;
      if (BITS == 2) s = s == 0 ? 0 : (s == 5 ? 3 : s - 1);
      key = (key << BITS) | s;
    }
    if (BITS * SPK < 64) key <<= (64 - BITS * SPK);
    keys[t] = key;
    vals[t] = uint32_t(i);
  }
}

__global__ void flag_ties_kernel(const uint64_t* __restrict__ keys, uint64_t n1, uint8_t* __restrict__ tied) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n1; i += stride) {
    bool t = (i > 0 && keys[i] == keys[i - 1]) || (i + 1 < n1 && keys[i] == keys[i + 1]);
    tied[i] = t ? 1 : 0;
  }
}

// One thread block per 256-row reference block: planes via warp ballots, per-block symbol counts.
template <int ALPHA, class SaT>
__global__ void __launch_bounds__(256)
    bwt_blocks_kernel(const uint8_t* __restrict__ sym, const SaT* __restrict__ sa, uint64_t n1,
                      uint64_t* __restrict__ blocks, uint32_t* __restrict__ counts /* [card][n_blocks] */,
                      uint64_t n_blocks) {
  constexpr int PLANES = ALPHA == 0 ? 3 : 5;
  constexpr int CARD = ALPHA == 0 ? 6 : 22;
  constexpr int WORDS = PLANES * 4 + (ALPHA == 0 ? 8 : 24);
  __shared__ uint32_t plane_words[PLANES][8];
  __shared__ uint32_t cnt[CARD];
  uint64_t b = blockIdx.x;
  uint64_t row = b * 256 + threadIdx.x;
  uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < CARD) cnt[threadIdx.x] = 0;
  __syncthreads();
  bool valid = row < n1;
  uint8_t idx = 0, code = 0;
  if (valid) {
    const SaT v = sa[row];
    idx = v == 0 ? 0 : sym[v - 1];  // fm_index.rs:220-223
    code = index_to_code(ALPHA, idx);
  }
#pragma unroll
  for (int p = 0; p < PLANES; p++) {
    uint32_t w = __ballot_sync(0xffffffffu, valid && ((code >> p) & 1));
    if (lane == 0) plane_words[p][warp] = w;
  }
  for (int s = 0; s < CARD; s++) {
    uint32_t m = __ballot_sync(0xffffffffu, valid && idx == s);
    if (lane == 0 && m) atomicAdd(&cnt[s], __popc(m));
  }
  __syncthreads();
  uint64_t* out = blocks + b * WORDS;
  if (threadIdx.x < PLANES * 4) {
    int p = threadIdx.x / 4, w = threadIdx.x % 4;
    out[threadIdx.x] = uint64_t(plane_words[p][2 * w]) | (uint64_t(plane_words[p][2 * w + 1]) << 32);
  }
  if (threadIdx.x < CARD) counts[uint64_t(threadIdx.x) * n_blocks + b] = cnt[threadIdx.x];
}

template <int ALPHA>
__global__ void write_milestones_kernel(const uint64_t* __restrict__ ms /* [card][n_blocks] */,
                                        uint64_t n_blocks, uint64_t* __restrict__ blocks) {
  constexpr int PLANES = ALPHA == 0 ? 3 : 5;
  constexpr int CARD = ALPHA == 0 ? 6 : 22;
  constexpr int NMS = ALPHA == 0 ? 8 : 24;
  constexpr int WORDS = PLANES * 4 + NMS;
  uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (t >= n_blocks * NMS) return;
  uint64_t b = t / NMS;
  int s = int(t % NMS);
  blocks[b * WORDS + PLANES * 4 + s] = s < CARD ? ms[uint64_t(s) * n_blocks + b] : 0;
}

// one thread per output word of the bit-packed sampled SA
template <class SaT>
__global__ void pack_sa_kernel(const SaT* __restrict__ sa, uint64_t n1, uint64_t ratio, uint32_t bits,
                               uint64_t n_words, uint64_t* __restrict__ words) {
  uint64_t w = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (w >= n_words) return;
  uint64_t n_elems = (n1 + ratio - 1) / ratio;
  uint64_t lo_bit = w * 64, hi_bit = lo_bit + 63;
  uint64_t e0 = lo_bit / bits, e1 = hi_bit / bits;
  uint64_t out = 0;
  for (uint64_t e = e0; e <= e1 && e < n_elems; e++) {
    uint64_t v = sa[e * ratio];
    uint64_t ebit = e * bits;
    if (ebit >= lo_bit)
      out |= v << (ebit - lo_bit);
    else
      out |= v >> (lo_bit - ebit);
  }
  words[w] = out;
}

// ---- prefix doubling on the GPU (for repeat-rich text where many suffixes share the packed key) ----
struct MaxOp {
  __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
__global__ void pd_heads_kernel(const uint64_t* __restrict__ keys, uint64_t n1, uint8_t* __restrict__ head,
                                uint32_t* __restrict__ v) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n1; i += stride) {
    bool hd = i == 0 || keys[i] != keys[i - 1];
    head[i] = hd ? 1 : 0;
    v[i] = hd ? uint32_t(i) : 0u;
  }
}
__global__ void pd_scatter_rank_kernel(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ grp, uint64_t n1,
                                       uint32_t* __restrict__ rank) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n1; i += stride) rank[sa[i]] = grp[i];
}
__global__ void pd_flag_unresolved_kernel(const uint8_t* __restrict__ head, uint64_t n1, uint8_t* __restrict__ unres) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n1; i += stride)
    unres[i] = (head[i] && (i + 1 == n1 || head[i + 1])) ? 0 : 1;
}
__global__ void pd_keys2_kernel(const uint32_t* __restrict__ U, uint64_t m, const uint32_t* __restrict__ sa,
                                const uint32_t* __restrict__ grp, const uint32_t* __restrict__ rank, uint64_t h,
                                uint64_t n1, uint64_t* __restrict__ keys2, uint32_t* __restrict__ vals2) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; j < m; j += stride) {
    uint32_t slot = U[j], pos = sa[slot];
    uint64_t nxt = uint64_t(pos) + h;
    uint32_t k2 = nxt < n1 ? rank[nxt] + 1u : 0u;  // a suffix that ends first sorts first
    keys2[j] = (uint64_t(grp[slot]) << 32) | k2;
    vals2[j] = pos;
  }
}
__global__ void pd_apply_kernel(const uint32_t* __restrict__ U, uint64_t m, const uint64_t* __restrict__ keys2,
                                const uint32_t* __restrict__ vals2, uint32_t* __restrict__ sa,
                                uint8_t* __restrict__ head, uint32_t* __restrict__ v) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; j < m; j += stride) {
    uint32_t slot = U[j];
    bool hd = j == 0 || keys2[j] != keys2[j - 1];
    sa[slot] = vals2[j];
    head[slot] = hd ? 1 : 0;
    v[j] = hd ? slot : 0u;
  }
}
__global__ void pd_regroup_kernel(const uint32_t* __restrict__ U, uint64_t m, const uint32_t* __restrict__ g2,
                                  const uint32_t* __restrict__ sa, uint32_t* __restrict__ grp,
                                  uint32_t* __restrict__ rank) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t j = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; j < m; j += stride) {
    uint32_t slot = U[j];
    grp[slot] = g2[j];
    rank[sa[slot]] = g2[j];
  }
}

// keys_sorted / sa: result of the first radix sort on keys of h0 symbols in which equal keys imply
// equal first h0 symbols (3- or 5-bit keys).  Refines sa in place to the full suffix order.
// kbuf[2]: two 8*n1-byte scratch buffers (the radix sort's key buffers), vscratch: 4*n1 bytes.
int prefix_doubling(const uint64_t* keys_sorted, uint32_t* sa, uint64_t n1, uint64_t h0, uint64_t* kbuf0,
                    uint64_t* kbuf1, uint32_t* vscratch, int* rounds_out) {
  uint32_t *grp = nullptr, *rank = nullptr, *U = nullptr, *v = nullptr, *vals_out = nullptr;
  uint8_t *head = nullptr, *unres = nullptr;
  unsigned long long* d_num = nullptr;
  void* d_temp = nullptr;
  size_t temp_cap = 0;
  auto need_temp = [&](size_t b) {
    if (b > temp_cap) {
      cudaFree(d_temp);
      CU(cudaMalloc(&d_temp, b + 256));
      temp_cap = b + 256;
    }
  };
  int rounds = 0;
  try {
    CU(cudaMalloc(&grp, n1 * 4));
    CU(cudaMalloc(&rank, n1 * 4));
    CU(cudaMalloc(&U, n1 * 4));
    CU(cudaMalloc(&v, n1 * 4));
    CU(cudaMalloc(&vals_out, n1 * 4));
    CU(cudaMalloc(&head, n1));
    CU(cudaMalloc(&unres, n1));
    CU(cudaMalloc(&d_num, 8));
    const unsigned G = 148 * 16;
    pd_heads_kernel<<<G, 256>>>(keys_sorted, n1, head, v);
    size_t tb = 0;
    CU(cub::DeviceScan::InclusiveScan(nullptr, tb, v, grp, MaxOp(), (long long)n1));
    need_temp(tb);
    CU(cub::DeviceScan::InclusiveScan(d_temp, tb, v, grp, MaxOp(), (long long)n1));
    pd_scatter_rank_kernel<<<G, 256>>>(sa, grp, n1, rank);
    // from here on the key buffers are scratch
    for (uint64_t h = h0;; h *= 2) {
      pd_flag_unresolved_kernel<<<G, 256>>>(head, n1, unres);
      thrust::counting_iterator<uint32_t> it(0);
      CU(cub::DeviceSelect::Flagged(nullptr, tb, it, unres, U, d_num, (long long)n1));
      need_temp(tb);
      CU(cub::DeviceSelect::Flagged(d_temp, tb, it, unres, U, d_num, (long long)n1));
      unsigned long long m = 0;
      CU(cudaMemcpy(&m, d_num, 8, cudaMemcpyDeviceToHost));
      if (m == 0) break;
      if (h >= 2 * n1) throw std::string("prefix doubling did not converge");
      rounds++;
      pd_keys2_kernel<<<G, 256>>>(U, m, sa, grp, rank, h, n1, kbuf0, vscratch);
      cub::DoubleBuffer<uint64_t> kb(kbuf0, kbuf1);
      cub::DoubleBuffer<uint32_t> vb(vscratch, vals_out);
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (long long)m, 0, 64));
      need_temp(tb);
      CU(cub::DeviceRadixSort::SortPairs(d_temp, tb, kb, vb, (long long)m, 0, 64));
      pd_apply_kernel<<<G, 256>>>(U, m, kb.Current(), vb.Current(), sa, head, v);
      uint32_t* g2 = vb.Alternate();  // scratch for the regrouped starts
      CU(cub::DeviceScan::InclusiveScan(nullptr, tb, v, g2, MaxOp(), (long long)m));
      need_temp(tb);
      CU(cub::DeviceScan::InclusiveScan(d_temp, tb, v, g2, MaxOp(), (long long)m));
      pd_regroup_kernel<<<G, 256>>>(U, m, g2, sa, grp, rank);
      CU(cudaDeviceSynchronize());
    }
  } catch (const std::string& msg) {
    cudaFree(grp); cudaFree(rank); cudaFree(U); cudaFree(v); cudaFree(vals_out); cudaFree(head); cudaFree(unres);
    cudaFree(d_num); cudaFree(d_temp);
    throw;
  }
  cudaFree(grp); cudaFree(rank); cudaFree(U); cudaFree(v); cudaFree(vals_out); cudaFree(head); cudaFree(unres);
  cudaFree(d_num); cudaFree(d_temp);
  if (rounds_out) *rounds_out = rounds;
  return 0;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

unsigned bits_per_element(uint64_t bwt_len) {
  uint64_t v = bwt_len - 1;
  return v ? 64u - unsigned(__builtin_clzll(v)) : 0u;
}

template <int BITS>
void make_keys(const uint8_t* sym, uint64_t n1, uint64_t* keys, uint32_t* vals) {
  make_keys_kernel<BITS><<<148 * 16, 256>>>(sym, n1, 0, n1, keys, vals);
}



// Builds the reference-layout parts of text[0..n) + '$' (ASCII, host or device pointer).
int build_parts_impl(int alphabet, const uint8_t* text, bool text_on_device, uint64_t n, uint64_t ratio, int device,
                     uint64_t* blocks_out, uint64_t* prefix_sums_out, uint64_t* sa_words_out, double* phase_s,
                     std::string& err, awry::DeviceParts* keep) {
  uint8_t* d_sym = nullptr;
  uint64_t *d_keys[2] = {nullptr, nullptr}, *d_blocks = nullptr, *d_ms = nullptr, *d_saw = nullptr;
  uint32_t *d_vals[2] = {nullptr, nullptr}, *d_counts = nullptr;
  uint8_t* d_tied = nullptr;
  void* d_temp = nullptr;
  int rc = 0;
  double ph[8] = {0};
  try {
    if (ratio == 0) throw std::string("ratio must be >= 1");
    const uint64_t n1 = n + 1;
    if (n1 >= (1ull << 32) - 256) throw std::string("internal: 2^32 symbols or more belong to build_parts_wide_impl");
    CU(cudaSetDevice(device));
    const int card = alphabet == 0 ? 6 : 22;
    const size_t words_per_block = alphabet == 0 ? 20 : 44;
    const uint64_t n_blocks = (n1 + 255) / 256;
    double t0 = now_s(), t = t0;
    CU(cudaMalloc(&d_sym, n1 + 64));
    bool pure2 = alphabet == 0;
    {
      CU(cudaMemcpy(d_sym, text, n, text_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
      unsigned int* d_flags = nullptr;
      CU(cudaMalloc(&d_flags, 8));
      CU(cudaMemset(d_flags, 0, 8));
      ascii_to_sym_kernel<<<148 * 16, 256>>>(alphabet, n, d_sym, d_flags);
      unsigned int flags[2] = {0, 0};
      CU(cudaMemcpy(flags, d_flags, 8, cudaMemcpyDeviceToHost));
      cudaFree(d_flags);
      if (flags[0]) throw std::string("text contains a sentinel ('$' or '#')");
      if (flags[1]) pure2 = false;
    }
    CU(cudaDeviceSynchronize());
    ph[0] = now_s() - t;
    t = now_s();

    for (int i = 0; i < 2; i++) {
      CU(cudaMalloc(&d_keys[i], n1 * 8));
      CU(cudaMalloc(&d_vals[i], n1 * 4));
    }
    if (alphabet == 0 && pure2)
      make_keys<2>(d_sym, n1, d_keys[0], d_vals[0]);
    else if (alphabet == 0)
      make_keys<3>(d_sym, n1, d_keys[0], d_vals[0]);
    else
      make_keys<5>(d_sym, n1, d_keys[0], d_vals[0]);
    CU(cudaDeviceSynchronize());
    ph[1] = now_s() - t;
    t = now_s();

    cub::DoubleBuffer<uint64_t> kb(d_keys[0], d_keys[1]);
    cub::DoubleBuffer<uint32_t> vb(d_vals[0], d_vals[1]);
    size_t temp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, (long long)n1, 0, 64));
    CU(cudaMalloc(&d_temp, temp_bytes));
    CU(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, (long long)n1, 0, 64));
    CU(cudaDeviceSynchronize());
    cudaFree(d_temp);
    d_temp = nullptr;
    ph[2] = now_s() - t;
    t = now_s();
    uint64_t* keys_sorted = kb.Current();
    uint32_t* sa = vb.Current();

    // ---- ties: rows whose key equals a neighbour's; fixed on the host by suffix comparison
    CU(cudaMalloc(&d_tied, n1));
    flag_ties_kernel<<<148 * 16, 256>>>(keys_sorted, n1, d_tied);
    // compact the tied row numbers
    uint32_t* d_rows = vb.Alternate();  // scratch: the other value buffer
    unsigned long long* d_num = nullptr;
    CU(cudaMalloc(&d_num, 8));
    {
      thrust::counting_iterator<uint32_t> it(0);
      size_t tb = 0;
      CU(cub::DeviceSelect::Flagged(nullptr, tb, it, d_tied, d_rows, d_num, (long long)n1));
      CU(cudaMalloc(&d_temp, tb));
      CU(cub::DeviceSelect::Flagged(d_temp, tb, it, d_tied, d_rows, d_num, (long long)n1));
      CU(cudaDeviceSynchronize());
      cudaFree(d_temp);
      d_temp = nullptr;
    }
    unsigned long long n_tied = 0;
    CU(cudaMemcpy(&n_tied, d_num, 8, cudaMemcpyDeviceToHost));
    cudaFree(d_num);
    unsigned long long HOST_FIXUP_MAX = 2000000ull;
    if (const char* e = getenv("AWRY_B200_BUILD_HOST_FIXUP_MAX")) HOST_FIXUP_MAX = strtoull(e, nullptr, 10);
    if (n_tied > HOST_FIXUP_MAX) {
      // repeat-rich text: GPU prefix doubling.  It needs keys whose equality implies equality of the
      // first h0 symbols, which the 2-bit keys (sentinel packed as 'A') do not give: redo with 3 bits.
      cudaFree(d_tied);
      d_tied = nullptr;
      uint64_t h0 = alphabet == 0 ? 21 : 12;
      if (alphabet == 0 && pure2) {
        make_keys<3>(d_sym, n1, d_keys[0], d_vals[0]);
        cub::DoubleBuffer<uint64_t> kb3(d_keys[0], d_keys[1]);
        cub::DoubleBuffer<uint32_t> vb3(d_vals[0], d_vals[1]);
        size_t tb3 = 0;
        CU(cub::DeviceRadixSort::SortPairs(nullptr, tb3, kb3, vb3, (long long)n1, 0, 64));
        CU(cudaMalloc(&d_temp, tb3));
        CU(cub::DeviceRadixSort::SortPairs(d_temp, tb3, kb3, vb3, (long long)n1, 0, 64));
        CU(cudaDeviceSynchronize());
        cudaFree(d_temp);
        d_temp = nullptr;
        keys_sorted = kb3.Current();
        sa = vb3.Current();
      }
      uint64_t* other_keys = keys_sorted == d_keys[0] ? d_keys[1] : d_keys[0];
      uint32_t* other_vals = sa == d_vals[0] ? d_vals[1] : d_vals[0];
      // the doubling rounds reuse the key buffers as scratch once the group heads are extracted;
      // keys_sorted itself is only read by the first kernel, so it may serve as scratch too
      int rounds = 0;
      prefix_doubling(keys_sorted, sa, n1, h0, other_keys, const_cast<uint64_t*>(keys_sorted), other_vals, &rounds);
      ph[3] = now_s() - t;
      t = now_s();
    } else if (n_tied > 0) {
      std::vector<uint32_t> rows(n_tied), pos(n_tied);
      std::vector<uint64_t> rkeys(n_tied);
      CU(cudaMemcpy(rows.data(), d_rows, n_tied * 4, cudaMemcpyDeviceToHost));
      // gather positions and keys of the tied rows
      for (unsigned long long i = 0; i < n_tied;) {  // runs of consecutive rows -> few large copies
        unsigned long long j = i + 1;
        while (j < n_tied && rows[j] == rows[j - 1] + 1) j++;
        CU(cudaMemcpy(pos.data() + i, sa + rows[i], (j - i) * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(rkeys.data() + i, keys_sorted + rows[i], (j - i) * 8, cudaMemcpyDeviceToHost));
        i = j;
      }
      std::vector<uint8_t> hsym(n1);
      CU(cudaMemcpy(hsym.data(), d_sym, n1, cudaMemcpyDeviceToHost));
      auto less = [&](uint32_t a, uint32_t b) {
        if (a == b) return false;
        uint64_t k = 0;
        while (hsym[a + k] == hsym[b + k]) k++;  // terminates: '$' is unique
        return hsym[a + k] < hsym[b + k];
      };
      for (unsigned long long i = 0; i < n_tied;) {
        unsigned long long j = i + 1;
        while (j < n_tied && rows[j] == rows[j - 1] + 1 && rkeys[j] == rkeys[i]) j++;
        std::sort(pos.begin() + i, pos.begin() + j, less);
        CU(cudaMemcpy(sa + rows[i], pos.data() + i, (j - i) * 4, cudaMemcpyHostToDevice));
        i = j;
      }
    }
    cudaFree(d_tied);
    d_tied = nullptr;
    if (n_tied <= HOST_FIXUP_MAX) {
      ph[3] = now_s() - t;
      t = now_s();
    }

    // keys are no longer needed: free them before allocating the outputs
    uint32_t* sa_keep = sa;
    for (int i = 0; i < 2; i++) {
      cudaFree(d_keys[i]);
      d_keys[i] = nullptr;
      if (d_vals[i] != sa_keep) {
        cudaFree(d_vals[i]);
        d_vals[i] = nullptr;
      }
    }

    CU(cudaMalloc(&d_blocks, n_blocks * words_per_block * 8));
    CU(cudaMemset(d_blocks, 0, n_blocks * words_per_block * 8));
    CU(cudaMalloc(&d_counts, uint64_t(card) * n_blocks * 4));
    CU(cudaMalloc(&d_ms, uint64_t(card) * n_blocks * 8));
    if (alphabet == 0)
      bwt_blocks_kernel<0, uint32_t><<<unsigned(n_blocks), 256>>>(d_sym, sa, n1, d_blocks, d_counts, n_blocks);
    else
      bwt_blocks_kernel<1, uint32_t><<<unsigned(n_blocks), 256>>>(d_sym, sa, n1, d_blocks, d_counts, n_blocks);
    CU(cudaDeviceSynchronize());
    ph[4] = now_s() - t;
    t = now_s();

    {
      size_t tb = 0;
      CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_counts, d_ms, (long long)n_blocks));
      CU(cudaMalloc(&d_temp, tb));
      std::vector<uint64_t> totals(card, 0);
      for (int s = 0; s < card; s++) {
        CU(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts + uint64_t(s) * n_blocks,
                                         d_ms + uint64_t(s) * n_blocks, (long long)n_blocks));
        uint64_t last_ms = 0;
        uint32_t last_cnt = 0;
        CU(cudaMemcpy(&last_ms, d_ms + uint64_t(s) * n_blocks + n_blocks - 1, 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&last_cnt, d_counts + uint64_t(s) * n_blocks + n_blocks - 1, 4, cudaMemcpyDeviceToHost));
        totals[s] = last_ms + last_cnt;
      }
      cudaFree(d_temp);
      d_temp = nullptr;
      uint64_t acc = 0;  // fm_index.rs:233-240
      for (int c = 0; c <= card; c++) {
        prefix_sums_out[c] = acc;
        if (c < card) acc += totals[c];
      }
      const int nms = alphabet == 0 ? 8 : 24;
      uint64_t threads = n_blocks * nms;
      if (alphabet == 0)
        write_milestones_kernel<0><<<unsigned((threads + 255) / 256), 256>>>(d_ms, n_blocks, d_blocks);
      else
        write_milestones_kernel<1><<<unsigned((threads + 255) / 256), 256>>>(d_ms, n_blocks, d_blocks);
      CU(cudaDeviceSynchronize());
    }
    cudaFree(d_counts);
    d_counts = nullptr;
    cudaFree(d_ms);
    d_ms = nullptr;
    ph[5] = now_s() - t;
    t = now_s();

    unsigned bits = bits_per_element(n1);
    uint64_t n_elems = (n1 + ratio - 1) / ratio;
    uint64_t n_words = uint64_t(((unsigned __int128)n_elems * bits + 63) / 64);
    CU(cudaMalloc(&d_saw, (n_words + 1) * 8));
    pack_sa_kernel<uint32_t><<<unsigned((n_words + 255) / 256), 256>>>(sa, n1, ratio, bits, n_words, d_saw);
    CU(cudaDeviceSynchronize());
    if (sa_words_out) CU(cudaMemcpy(sa_words_out, d_saw, n_words * 8, cudaMemcpyDeviceToHost));
    if (blocks_out) CU(cudaMemcpy(blocks_out, d_blocks, n_blocks * words_per_block * 8, cudaMemcpyDeviceToHost));
    if (keep) {
      keep->d_blocks = d_blocks;
      keep->d_sa_words = d_saw;
      keep->n_block_words = n_blocks * words_per_block;
      keep->n_sa_words = n_words;
      keep->device = device;
      d_blocks = nullptr;
      d_saw = nullptr;
    }
    ph[6] = now_s() - t;
    ph[7] = now_s() - t0;
  } catch (const std::string& m) {
    err = m;
    rc = -1;
  }
  cudaFree(d_sym);
  for (int i = 0; i < 2; i++) {
    cudaFree(d_keys[i]);
    cudaFree(d_vals[i]);
  }
  cudaFree(d_blocks);
  cudaFree(d_ms);
  cudaFree(d_saw);
  cudaFree(d_counts);
  cudaFree(d_tied);
  cudaFree(d_temp);
  if (phase_s) memcpy(phase_s, ph, sizeof ph);
  return rc;
}

// ---- texts of 2^32 symbols or more: 64-bit suffix-array elements, sorted bucket by bucket ----
// One radix sort of (64-bit key, 64-bit position) pairs over 5 G suffixes needs 160 GB of buffers.  The
// suffixes are therefore sorted per FIRST SYMBOL (6 / 22 buckets, already in suffix order between them):
// positions of the bucket selected with a stream compaction, keys of the following symbols packed as above,
// one radix sort per bucket, written behind the previous bucket.  Ties (suffixes sharing the whole packed key:
// a handful on random text) are fixed on the host by suffix comparison; repeat-rich text beyond 2^32 symbols
// is refused (the prefix-doubling rounds above keep 32-bit ranks).  Also runs on small texts when
// AWRY_B200_BUILD_WIDE=1 (tests).
struct FirstSymbolIs {
  const uint8_t* sym;
  uint8_t want;
  __host__ __device__ bool operator()(uint64_t i) const { return sym[i] == want; }
};
struct FirstSymbolCount {  // the same as a 0/1 summand
  const uint8_t* sym;
  uint8_t want;
  __host__ __device__ unsigned long long operator()(uint64_t i) const { return sym[i] == want ? 1ull : 0ull; }
};

template <int BITS>
__global__ void make_keys_at_kernel(const uint8_t* __restrict__ sym, uint64_t n1, const uint64_t* __restrict__ pos,
                                    uint64_t count, uint64_t* __restrict__ keys) {
  constexpr int SPK = 64 / BITS;
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < count; t += stride) {
    const uint64_t i = pos[t] + 1;  // the first symbol is the bucket's: keys start behind it
    uint64_t key = 0;
#pragma unroll 8
    for (int j = 0; j < SPK; j++) {
      uint32_t s = i + j < n1 ? sym[i + j] : 0;
      if (BITS == 2) s = s == 0 ? 0 : (s == 5 ? 3 : s - 1);
      key = (key << BITS) | s;
    }
    if (BITS * SPK < 64) key <<= (64 - BITS * SPK);
    keys[t] = key;
  }
}

int build_parts_wide_impl(int alphabet, const uint8_t* text, bool text_on_device, uint64_t n, uint64_t ratio, int device,
                          uint64_t* blocks_out, uint64_t* prefix_sums_out, uint64_t* sa_words_out, double* phase_s,
                          std::string& err, awry::DeviceParts* keep) {
  uint8_t* d_sym = nullptr;
  uint64_t *d_sa = nullptr, *d_keys[2] = {nullptr, nullptr}, *d_pos[2] = {nullptr, nullptr};
  uint64_t *d_blocks = nullptr, *d_ms = nullptr, *d_saw = nullptr;
  uint32_t* d_counts = nullptr;
  uint8_t* d_tied = nullptr;
  void* d_temp = nullptr;
  unsigned long long* d_num = nullptr;
  int rc = 0;
  double ph[8] = {0};
  try {
    if (ratio == 0) throw std::string("ratio must be >= 1");
    const uint64_t n1 = n + 1;
    CU(cudaSetDevice(device));
    const int card = alphabet == 0 ? 6 : 22;
    const size_t words_per_block = alphabet == 0 ? 20 : 44;
    const uint64_t n_blocks = (n1 + 255) / 256;
    double t0 = now_s(), t = t0;
    CU(cudaMalloc(&d_sym, n1 + 64));
    bool pure2 = alphabet == 0;
    {
      CU(cudaMemcpy(d_sym, text, n, text_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
      unsigned int* d_flags = nullptr;
      CU(cudaMalloc(&d_flags, 8));
      CU(cudaMemset(d_flags, 0, 8));
      ascii_to_sym_kernel<<<148 * 16, 256>>>(alphabet, n, d_sym, d_flags);
      unsigned int flags[2] = {0, 0};
      CU(cudaMemcpy(flags, d_flags, 8, cudaMemcpyDeviceToHost));
      cudaFree(d_flags);
      if (flags[0]) throw std::string("text contains a sentinel ('$' or '#')");
      if (flags[1]) pure2 = false;
    }
    ph[0] = now_s() - t;
    t = now_s();

    // symbol histogram on the host side of a device reduction: bucket sizes
    std::vector<uint64_t> bucket(card, 0);
    CU(cudaMalloc(&d_num, 8));
    CU(cudaMalloc(&d_sa, n1 * 8));
    uint64_t max_bucket = 0;
    {
      // sizes first (one counting pass per symbol is cheap next to the sorts)
      for (int s = 0; s < card; s++) {
        thrust::counting_iterator<uint64_t> it(0);
        auto flag = thrust::make_transform_iterator(it, FirstSymbolCount{d_sym, uint8_t(s)});
        size_t tb = 0;
        CU(cub::DeviceReduce::Sum(nullptr, tb, flag, d_num, (long long)n1));
        CU(cudaMalloc(&d_temp, tb + 16));
        CU(cub::DeviceReduce::Sum(d_temp, tb, flag, d_num, (long long)n1));
        unsigned long long c = 0;
        CU(cudaMemcpy(&c, d_num, 8, cudaMemcpyDeviceToHost));
        cudaFree(d_temp);
        d_temp = nullptr;
        bucket[s] = c;
        max_bucket = std::max<uint64_t>(max_bucket, c);
      }
    }
    for (int i = 0; i < 2; i++) {
      CU(cudaMalloc(&d_keys[i], (max_bucket + 1) * 8));
      CU(cudaMalloc(&d_pos[i], (max_bucket + 1) * 8));
    }
    CU(cudaMalloc(&d_tied, max_bucket + 1));
    ph[1] = now_s() - t;
    t = now_s();

    std::vector<uint8_t> hsym;  // host copy of the symbols, fetched only if a tie needs it
    uint64_t base = 0, ties_total = 0;
    for (int s = 0; s < card; s++) {
      const uint64_t cnt = bucket[s];
      if (cnt == 0) continue;
      {  // positions whose first symbol is s, ascending
        thrust::counting_iterator<uint64_t> it(0);
        size_t tb = 0;
        CU(cub::DeviceSelect::If(nullptr, tb, it, d_pos[0], d_num, (long long)n1, FirstSymbolIs{d_sym, uint8_t(s)}));
        CU(cudaMalloc(&d_temp, tb + 16));
        CU(cub::DeviceSelect::If(d_temp, tb, it, d_pos[0], d_num, (long long)n1, FirstSymbolIs{d_sym, uint8_t(s)}));
        cudaFree(d_temp);
        d_temp = nullptr;
      }
      if (alphabet == 0 && pure2)
        make_keys_at_kernel<2><<<148 * 16, 256>>>(d_sym, n1, d_pos[0], cnt, d_keys[0]);
      else if (alphabet == 0)
        make_keys_at_kernel<3><<<148 * 16, 256>>>(d_sym, n1, d_pos[0], cnt, d_keys[0]);
      else
        make_keys_at_kernel<5><<<148 * 16, 256>>>(d_sym, n1, d_pos[0], cnt, d_keys[0]);
      cub::DoubleBuffer<uint64_t> kb(d_keys[0], d_keys[1]);
      cub::DoubleBuffer<uint64_t> vb(d_pos[0], d_pos[1]);
      size_t tb = 0;
      CU(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, (long long)cnt, 0, 64));
      CU(cudaMalloc(&d_temp, tb + 16));
      CU(cub::DeviceRadixSort::SortPairs(d_temp, tb, kb, vb, (long long)cnt, 0, 64));
      CU(cudaDeviceSynchronize());
      cudaFree(d_temp);
      d_temp = nullptr;
      uint64_t* keys_sorted = kb.Current();
      uint64_t* sa_b = vb.Current();
      // ties inside the bucket
      flag_ties_kernel<<<148 * 16, 256>>>(keys_sorted, cnt, d_tied);
      uint64_t* d_rows = vb.Alternate();
      {
        thrust::counting_iterator<uint64_t> it(0);
        size_t tb2 = 0;
        CU(cub::DeviceSelect::Flagged(nullptr, tb2, it, d_tied, d_rows, d_num, (long long)cnt));
        CU(cudaMalloc(&d_temp, tb2 + 16));
        CU(cub::DeviceSelect::Flagged(d_temp, tb2, it, d_tied, d_rows, d_num, (long long)cnt));
        CU(cudaDeviceSynchronize());
        cudaFree(d_temp);
        d_temp = nullptr;
      }
      unsigned long long n_tied = 0;
      CU(cudaMemcpy(&n_tied, d_num, 8, cudaMemcpyDeviceToHost));
      ties_total += n_tied;
      if (ties_total > 4000000ull)
        throw std::string("repeat-rich text of 2^32 symbols or more: too many suffixes share their first 21-32 "
                          "symbols for the host fix-up, and the prefix-doubling rounds keep 32-bit ranks");
      if (n_tied > 0) {
        std::vector<uint64_t> rows(n_tied), pos(n_tied), rkeys(n_tied);
        CU(cudaMemcpy(rows.data(), d_rows, n_tied * 8, cudaMemcpyDeviceToHost));
        for (unsigned long long i = 0; i < n_tied;) {  // runs of consecutive rows -> few large copies
          unsigned long long j = i + 1;
          while (j < n_tied && rows[j] == rows[j - 1] + 1) j++;
          CU(cudaMemcpy(pos.data() + i, sa_b + rows[i], (j - i) * 8, cudaMemcpyDeviceToHost));
          CU(cudaMemcpy(rkeys.data() + i, keys_sorted + rows[i], (j - i) * 8, cudaMemcpyDeviceToHost));
          i = j;
        }
        if (hsym.empty()) {
          hsym.resize(n1);
          CU(cudaMemcpy(hsym.data(), d_sym, n1, cudaMemcpyDeviceToHost));
        }
        auto less = [&](uint64_t a, uint64_t b) {
          if (a == b) return false;
          uint64_t k = 0;
          while (hsym[a + k] == hsym[b + k]) k++;  // terminates: '$' is unique
          return hsym[a + k] < hsym[b + k];
        };
        for (unsigned long long i = 0; i < n_tied;) {
          unsigned long long j = i + 1;
          while (j < n_tied && rows[j] == rows[j - 1] + 1 && rkeys[j] == rkeys[i]) j++;
          std::sort(pos.begin() + i, pos.begin() + j, less);
          CU(cudaMemcpy(sa_b + rows[i], pos.data() + i, (j - i) * 8, cudaMemcpyHostToDevice));
          i = j;
        }
      }
      CU(cudaMemcpy(d_sa + base, sa_b, cnt * 8, cudaMemcpyDeviceToDevice));
      base += cnt;
    }
    if (base != n1) throw std::string("internal: buckets do not add up");
    for (int i = 0; i < 2; i++) {
      cudaFree(d_keys[i]);
      cudaFree(d_pos[i]);
      d_keys[i] = d_pos[i] = nullptr;
    }
    cudaFree(d_tied);
    d_tied = nullptr;
    ph[2] = now_s() - t;
    t = now_s();

    CU(cudaMalloc(&d_blocks, n_blocks * words_per_block * 8));
    CU(cudaMemset(d_blocks, 0, n_blocks * words_per_block * 8));
    CU(cudaMalloc(&d_counts, uint64_t(card) * n_blocks * 4));
    CU(cudaMalloc(&d_ms, uint64_t(card) * n_blocks * 8));
    if (alphabet == 0)
      bwt_blocks_kernel<0, uint64_t><<<unsigned(n_blocks), 256>>>(d_sym, d_sa, n1, d_blocks, d_counts, n_blocks);
    else
      bwt_blocks_kernel<1, uint64_t><<<unsigned(n_blocks), 256>>>(d_sym, d_sa, n1, d_blocks, d_counts, n_blocks);
    CU(cudaDeviceSynchronize());
    ph[4] = now_s() - t;
    t = now_s();
    {
      size_t tb = 0;
      CU(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_counts, d_ms, (long long)n_blocks));
      CU(cudaMalloc(&d_temp, tb + 16));
      std::vector<uint64_t> totals(card, 0);
      for (int s = 0; s < card; s++) {
        CU(cub::DeviceScan::ExclusiveSum(d_temp, tb, d_counts + uint64_t(s) * n_blocks, d_ms + uint64_t(s) * n_blocks,
                                         (long long)n_blocks));
        uint64_t last_ms = 0;
        uint32_t last_cnt = 0;
        CU(cudaMemcpy(&last_ms, d_ms + uint64_t(s) * n_blocks + n_blocks - 1, 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&last_cnt, d_counts + uint64_t(s) * n_blocks + n_blocks - 1, 4, cudaMemcpyDeviceToHost));
        totals[s] = last_ms + last_cnt;
      }
      cudaFree(d_temp);
      d_temp = nullptr;
      uint64_t acc = 0;  // fm_index.rs:233-240
      for (int c = 0; c <= card; c++) {
        prefix_sums_out[c] = acc;
        if (c < card) acc += totals[c];
      }
      const int nms = alphabet == 0 ? 8 : 24;
      const uint64_t threads = n_blocks * nms;
      if (alphabet == 0)
        write_milestones_kernel<0><<<unsigned((threads + 255) / 256), 256>>>(d_ms, n_blocks, d_blocks);
      else
        write_milestones_kernel<1><<<unsigned((threads + 255) / 256), 256>>>(d_ms, n_blocks, d_blocks);
      CU(cudaDeviceSynchronize());
    }
    cudaFree(d_counts);
    d_counts = nullptr;
    cudaFree(d_ms);
    d_ms = nullptr;
    ph[5] = now_s() - t;
    t = now_s();

    const unsigned bits = bits_per_element(n1);
    const uint64_t n_elems = (n1 + ratio - 1) / ratio;
    const uint64_t n_words = uint64_t(((unsigned __int128)n_elems * bits + 63) / 64);
    CU(cudaMalloc(&d_saw, (n_words + 1) * 8));
    pack_sa_kernel<uint64_t><<<unsigned((n_words + 255) / 256), 256>>>(d_sa, n1, ratio, bits, n_words, d_saw);
    CU(cudaDeviceSynchronize());
    cudaFree(d_sa);
    d_sa = nullptr;
    if (sa_words_out) CU(cudaMemcpy(sa_words_out, d_saw, n_words * 8, cudaMemcpyDeviceToHost));
    if (blocks_out) CU(cudaMemcpy(blocks_out, d_blocks, n_blocks * words_per_block * 8, cudaMemcpyDeviceToHost));
    if (keep) {
      keep->d_blocks = d_blocks;
      keep->d_sa_words = d_saw;
      keep->n_block_words = n_blocks * words_per_block;
      keep->n_sa_words = n_words;
      keep->device = device;
      d_blocks = nullptr;
      d_saw = nullptr;
    }
    ph[6] = now_s() - t;
    ph[7] = now_s() - t0;
  } catch (const std::string& m) {
    err = m;
    rc = -1;
  }
  cudaFree(d_sym);
  cudaFree(d_sa);
  for (int i = 0; i < 2; i++) {
    cudaFree(d_keys[i]);
    cudaFree(d_pos[i]);
  }
  cudaFree(d_blocks);
  cudaFree(d_ms);
  cudaFree(d_saw);
  cudaFree(d_counts);
  cudaFree(d_tied);
  cudaFree(d_temp);
  cudaFree(d_num);
  if (phase_s) memcpy(phase_s, ph, sizeof ph);
  return rc;
}



}  // namespace

namespace awry {

void DeviceParts::release() {
  if (!d_blocks && !d_sa_words) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  cudaFree(d_blocks);
  cudaFree(d_sa_words);
  d_blocks = d_sa_words = nullptr;
  if (prev >= 0) cudaSetDevice(prev);
}

int build_parts(int alphabet, const uint8_t* text, uint64_t n, uint64_t ratio, int device, uint64_t* blocks_out,
                uint64_t* prefix_sums_out, uint64_t* sa_words_out, double* phase_s, std::string& err,
                DeviceParts* keep) {
  if (alphabet != 0 && alphabet != 1) {
    err = "invalid alphabet";
    return -1;
  }
  if (n == 0 || text == nullptr) {
    err = "empty text";
    return -1;
  }
  cudaPointerAttributes a;
  bool on_device = cudaPointerGetAttributes(&a, text) == cudaSuccess && a.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  int prev = -1;
  cudaGetDevice(&prev);
  // 64-bit suffix-array elements from 2^32 - 256 symbols on (AWRY_B200_BUILD_WIDE=1: always -- tests)
  const char* bw = getenv("AWRY_B200_BUILD_WIDE");
  const bool wide = n + 1 >= (1ull << 32) - 256 || (bw && bw[0] == '1');
  int rc = wide ? build_parts_wide_impl(alphabet, text, on_device, n, ratio, device, blocks_out, prefix_sums_out,
                                        sa_words_out, phase_s, err, keep)
                : build_parts_impl(alphabet, text, on_device, n, ratio, device, blocks_out, prefix_sums_out, sa_words_out,
                                   phase_s, err, keep);
  if (prev >= 0) cudaSetDevice(prev);
  return rc;
}

// FASTA ('>') / FASTQ ('@') reader with the semantics the reference gets from
// libsufr::util::read_sequence_file (fm_index.rs:153): sequence lines of a record concatenated, records
// joined by `delimiter`, start offset and header text (without the marker) kept.  Case is left as it is:
// the device maps both cases to the same symbol (alphabet.rs:169-248), which is what upper-casing does.
// Large FASTA files are parsed by several threads, each on a byte range that begins at a line start
// (positional reads + memchr); FASTQ and small files take the sequential path.
namespace {

struct LocalParse {
  std::string text;
  std::vector<std::pair<uint64_t, std::string>> recs;  // (offset in `text`, header)
  std::string err;
};

// one FASTA line (no '\n'); `global_first`: nothing of the file precedes this parse
inline void fasta_line(const char* p, size_t n, char delimiter, bool global_first, LocalParse& out) {
  if (n && p[n - 1] == '\r') n--;
  if (n && p[0] == '>') {
    if (!(global_first && out.recs.empty() && out.text.empty())) out.text.push_back(delimiter);
    out.recs.emplace_back(out.text.size(), std::string(p + 1, n - 1));
    return;
  }
  if (!memchr(p, ' ', n) && !memchr(p, '\t', n)) {
    out.text.append(p, n);
    return;
  }
  for (size_t i = 0; i < n; i++)  // blanks inside sequence lines are dropped
    if (p[i] != ' ' && p[i] != '\t') out.text.push_back(p[i]);
}

// lines that START in [lo, hi) of the file
void fasta_range(int fd, uint64_t lo, uint64_t hi, uint64_t fsize, char delimiter, bool global_first, LocalParse& out) {
  std::vector<char> buf(8u << 20);
  std::string line;
  uint64_t pos = lo;
  bool done = false;
  out.text.reserve(size_t(hi - lo));
  while (!done && pos < fsize) {
    ssize_t got = pread(fd, buf.data(), buf.size(), off_t(pos));
    if (got <= 0) {
      out.err = "read error";
      return;
    }
    size_t off = 0;
    while (off < size_t(got)) {
      // a line that starts at or after `hi` belongs to the next range
      if (line.empty() && pos + off >= hi) {
        done = true;
        break;
      }
      const char* nl = static_cast<const char*>(memchr(buf.data() + off, '\n', size_t(got) - off));
      if (!nl) {
        line.append(buf.data() + off, size_t(got) - off);
        break;
      }
      size_t n = size_t(nl - (buf.data() + off));
      if (line.empty()) {
        fasta_line(buf.data() + off, n, delimiter, global_first, out);
      } else {
        line.append(buf.data() + off, n);
        fasta_line(line.data(), line.size(), delimiter, global_first, out);
        line.clear();
      }
      off += n + 1;
    }
    pos += uint64_t(got);
  }
  if (!line.empty()) fasta_line(line.data(), line.size(), delimiter, global_first, out);  // last line without '\n'
}

int finish_parse(std::vector<LocalParse>& parts, const std::string& path, awry::TextBuf& text,
                 std::vector<uint64_t>& starts, std::vector<std::string>& headers, std::string& err) {
  size_t total = 0, n_recs = 0;
  for (auto& lp : parts) {
    if (!lp.err.empty()) {
      err = lp.err + " in " + path;
      return -1;
    }
    total += lp.text.size();
    n_recs += lp.recs.size();
  }
  if (n_recs == 0) {
    err = "no sequence records in " + path;
    return -1;
  }
  if (total == 0) {
    err = "sequence records are empty in " + path;
    return -1;
  }
  text.p.reset(new uint8_t[total + 64]);
  text.n = total;
  std::vector<size_t> base(parts.size() + 1, 0);
  for (size_t i = 0; i < parts.size(); i++) base[i + 1] = base[i] + parts[i].text.size();
  std::vector<std::thread> th;
  for (size_t i = 0; i < parts.size(); i++)
    th.emplace_back([&, i] {
      if (!parts[i].text.empty()) memcpy(text.p.get() + base[i], parts[i].text.data(), parts[i].text.size());
      std::string().swap(parts[i].text);
    });
  for (auto& t : th) t.join();
  starts.reserve(n_recs);
  headers.reserve(n_recs);
  for (size_t i = 0; i < parts.size(); i++)
    for (auto& r : parts[i].recs) {
      starts.push_back(base[i] + r.first);
      headers.push_back(std::move(r.second));
    }
  return 0;
}

}  // namespace

int read_sequence_file(const std::string& path, char delimiter, TextBuf& text, std::vector<uint64_t>& starts,
                       std::vector<std::string>& headers, std::string& err) {
  text.reset();
  starts.clear();
  headers.clear();
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    err = "cannot open " + path;
    return -1;
  }
  struct Closer {
    int fd;
    ~Closer() { close(fd); }
  } closer{fd};
  {  // gzip-compressed input: inflate into an anonymous memory file, then parse that (same code below)
    unsigned char magic[2] = {0, 0};
    if (pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
      gzFile g = gzopen(path.c_str(), "rb");
      int mfd = memfd_create("awry_b200_inflated", 0);
      if (!g || mfd < 0) {
        if (g) gzclose(g);
        if (mfd >= 0) close(mfd);
        err = "cannot inflate " + path;
        return -1;
      }
      gzbuffer(g, 4u << 20);
      std::vector<char> zb(16u << 20);
      bool ok = true;
      for (;;) {
        int r = gzread(g, zb.data(), unsigned(zb.size()));
        if (r <= 0) {
          int zerr = Z_OK;
          gzerror(g, &zerr);
          if (r < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) ok = false;
          break;
        }
        for (int w = 0; w < r && ok;) {
          ssize_t k = write(mfd, zb.data() + w, size_t(r - w));
          if (k <= 0)
            ok = false;
          else
            w += int(k);
        }
        if (!ok) break;
      }
      gzclose(g);
      if (!ok) {
        close(mfd);
        err = "corrupt gzip stream or out of memory while inflating " + path;
        return -1;
      }
      close(closer.fd);
      closer.fd = fd = mfd;
    }
  }
  struct stat sb;
  if (fstat(fd, &sb) != 0) {
    err = "cannot stat " + path;
    return -1;
  }
  const uint64_t fsize = uint64_t(sb.st_size);
  // format: first non-blank byte
  uint64_t data_start = 0;
  char first = 0;
  {
    char head[4096];
    uint64_t pos = 0;
    while (!first && pos < fsize) {
      ssize_t got = pread(fd, head, sizeof head, off_t(pos));
      if (got <= 0) break;
      for (ssize_t i = 0; i < got; i++) {
        char c = head[i];
        if (c == '\n' || c == '\r') continue;
        first = c;
        data_start = pos + uint64_t(i);
        break;
      }
      pos += uint64_t(got);
    }
  }
  if (!first) {
    err = "no sequence records in " + path;
    return -1;
  }
  if (first != '>' && first != '@') {
    err = "input is neither FASTA ('>') nor FASTQ ('@')";
    return -1;
  }
  if (first == '>') {
    // ---- FASTA: ranges that begin at line starts, one thread each
    uint64_t range_min = 32u << 20;  // AWRY_B200_FASTA_RANGE_BYTES: smaller ranges (tests)
    if (const char* e = getenv("AWRY_B200_FASTA_RANGE_BYTES")) range_min = std::max<uint64_t>(16, strtoull(e, nullptr, 10));
    unsigned nt = unsigned(std::min<uint64_t>(std::min(8u, std::max(2u, std::thread::hardware_concurrency())),
                                              (fsize - data_start) / range_min + 1));
    std::vector<uint64_t> cut(nt + 1, fsize);
    cut[0] = data_start;
    bool ok = true;
    for (unsigned t = 1; t < nt && ok; t++) {
      uint64_t nominal = data_start + (fsize - data_start) * t / nt;
      // the first line start at or after `nominal`: just past the first '\n' at or after nominal - 1
      uint64_t pos = nominal - 1, found = fsize;
      char tmp[65536];
      for (int tries = 0; tries < 16 && found == fsize && pos < fsize; tries++) {  // up to 1 MiB
        ssize_t got = pread(fd, tmp, sizeof tmp, off_t(pos));
        if (got <= 0) break;
        const char* nl = static_cast<const char*>(memchr(tmp, '\n', size_t(got)));
        if (nl) found = pos + uint64_t(nl - tmp) + 1;
        pos += uint64_t(got);
      }
      if (found == fsize && pos < fsize) ok = false;  // a line longer than 1 MiB: one range after all
      cut[t] = std::max(found, cut[t - 1]);
    }
    if (!ok) {
      nt = 1;
      cut.assign({data_start, fsize});
    }
    std::vector<LocalParse> parts(nt);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t] { fasta_range(fd, cut[t], cut[t + 1], fsize, delimiter, t == 0, parts[t]); });
    for (auto& x : th) x.join();
    return finish_parse(parts, path, text, starts, headers, err);
  }
  // ---- FASTQ: four lines per record, sequential
  std::vector<LocalParse> parts(1);
  LocalParse& lp = parts[0];
  lp.text.reserve(size_t(fsize / 2));
  std::vector<char> buf(16u << 20);
  std::string line;
  int fq_line = 0;  // 0 header, 1 sequence, 2 '+', 3 qualities
  auto take_line = [&](const char* p, size_t n) {
    if (n && p[n - 1] == '\r') n--;
    if (fq_line == 0) {
      if (n == 0) return;  // blank lines between records
      if (!lp.recs.empty()) lp.text.push_back(delimiter);
      lp.recs.emplace_back(lp.text.size(), std::string(p + 1, n - 1));
    } else if (fq_line == 1) {
      lp.text.append(p, n);
    }
    fq_line = (fq_line + 1) & 3;
  };
  uint64_t pos = data_start;
  while (pos < fsize) {
    ssize_t got = pread(fd, buf.data(), buf.size(), off_t(pos));
    if (got <= 0) {
      err = "read error in " + path;
      return -1;
    }
    size_t off = 0;
    while (off < size_t(got)) {
      const char* nl = static_cast<const char*>(memchr(buf.data() + off, '\n', size_t(got) - off));
      if (!nl) {
        line.append(buf.data() + off, size_t(got) - off);
        break;
      }
      size_t n = size_t(nl - (buf.data() + off));
      if (line.empty()) {
        take_line(buf.data() + off, n);
      } else {
        line.append(buf.data() + off, n);
        take_line(line.data(), line.size());
        line.clear();
      }
      off += n + 1;
    }
    pos += uint64_t(got);
  }
  if (!line.empty()) take_line(line.data(), line.size());
  return finish_parse(parts, path, text, starts, headers, err);
}

}  // namespace awry
