// kernels_wide.cu -- the search path for indexes with bwt_len >= 2^32 - 256: 64-bit row pointers.
//
// The reference's SearchPtr is u64 (/root/reference/src/search.rs:7) and its sampled suffix array adapts its
// element width up to 64 bits (compressed_suffix_array.rs:124-130), so a drop-in must take any index the
// reference can build (wheat, pan-genomes, large protein collections).  Every BASELINE configuration fits 32-bit
// rows and runs the cooperative kernels of kernels.cu / kernels_locate.cu unchanged; an index past that limit
// keeps the SAME device blocks (their u32 counts become relative to 2^31-row superblocks, layout.cuh) and is
// searched by the kernels below: one thread per query / per hit, the plain algorithm of fm_index.rs:402-438,
// :516-544, :559-593 on 64-bit rows -- plus, for nucleotide search, a cooperative kernel (4 lanes per query, one
// line request per LF step).  No pair index, no unsampled array (both are 32-bit structures): a wide index is
// searched at the one-symbol-per-block rate.
#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "kernels_common.cuh"

namespace awry {

// ------------------------------------------------------------------ rank primitives on 64-bit rows

__device__ __forceinline__ uint64_t w_dna_occ(const WideView& ix, uint64_t pos, uint32_t c) {  // c in 0..4
  const uint64_t blk = pos >> 7;
  const uint32_t local = uint32_t(pos) & 127;
  AWRY_CHK(blk * DNA_BLOCK_UINT4 + 3 < ix.n_blocks_u4 && (blk >> (ix.sb_shift - 7)) < ix.n_sb);
  const DnaBlockRegs b = dna_load_block(ix.blocks, blk);
  const uint64_t* sb = ix.sb_counts + (blk >> (ix.sb_shift - 7)) * SB_STRIDE;
  uint64_t r;
  if (c < 4) {
    r = __ldg(sb + c) + (c == 0 ? b.ch[0].w : c == 1 ? b.ch[1].w : c == 2 ? b.ch[2].w : b.ch[3].w);
  } else {  // N: derived, as in layout.cuh
    const uint64_t start = blk << 7;
    const uint64_t acgt = __ldg(sb) + __ldg(sb + 1) + __ldg(sb + 2) + __ldg(sb + 3) + b.ch[0].w + b.ch[1].w + b.ch[2].w + b.ch[3].w;
    r = start - acgt - (ix.dollar_row < start ? 1u : 0u);
  }
  const uint32_t m0 = (c & 1) ? ~0u : 0u, m1 = (c & 2) ? ~0u : 0u, m2 = (c & 4) ? ~0u : 0u;
#pragma unroll
  for (int j = 0; j < 4; j++) r += __popc(dna_match(b.ch[j], m0, m1, m2) & chunk_mask(local, j));
  return r;
}
__device__ __forceinline__ uint32_t w_dna_code_at(const WideView& ix, uint64_t row) {
  AWRY_CHK((row >> 7) * DNA_BLOCK_UINT4 + 3 < ix.n_blocks_u4);
  const uint4 ch = ldg128(ix.blocks + size_t(row >> 7) * DNA_BLOCK_UINT4 + ((uint32_t(row) >> 5) & 3));
  const uint32_t t = uint32_t(row) & 31;
  return ((ch.x >> t) & 1u) | (((ch.y >> t) & 1u) << 1) | (((ch.z >> t) & 1u) << 2);
}
__device__ __forceinline__ uint64_t w_amino_occ(const WideView& ix, uint64_t pos, uint32_t s) {  // s in 1..21
  const uint64_t blk = pos >> 6;
  const uint32_t local = uint32_t(pos) & 63;
  AWRY_CHK(blk * AMINO_BLOCK_UINT4 + 7 < ix.n_blocks_u4 && (blk >> (ix.sb_shift - 6)) < ix.n_sb);
  const char* base = reinterpret_cast<const char*>(ix.blocks + size_t(blk) * AMINO_BLOCK_UINT4);
  uint64_t r = __ldg(ix.sb_counts + (blk >> (ix.sb_shift - 6)) * SB_STRIDE + s) +
               __ldg(reinterpret_cast<const uint32_t*>(base) + amino_count_word(s));
  const u32x8 c0 = ldg256(base);
  r += __popc(amino_match(c0, s) & chunk_mask(local, 0));
  if (local >= 32) {
    const u32x8 c1 = ldg256(base + 32);
    r += __popc(amino_match(c1, s) & chunk_mask(local, 1));
  }
  return r;
}
__device__ __forceinline__ uint32_t w_amino_symbol_at(const WideView& ix, uint64_t pos) {
  const uint32_t local = uint32_t(pos) & 63, j = local >> 5, t = local & 31;
  AWRY_CHK((pos >> 6) * AMINO_BLOCK_UINT4 + 7 < ix.n_blocks_u4);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(ix.blocks + size_t(pos >> 6) * AMINO_BLOCK_UINT4) + 8 * j;
  uint32_t s = 0;
#pragma unroll
  for (int p = 0; p < 5; p++) s |= ((__ldg(w + p) >> t) & 1u) << p;
  return s;
}

// update_range_with_symbol (fm_index.rs:559-582) for device symbol c; requires sp >= 1
template <int ALPHA>
__device__ __forceinline__ void w_lf_update(const WideView& ix, uint64_t& sp, uint64_t& ep, uint32_t c) {
  const uint64_t ra = ALPHA == 0 ? w_dna_occ(ix, sp - 1, c) : w_amino_occ(ix, sp - 1, c);
  const uint64_t rb = ALPHA == 0 ? w_dna_occ(ix, ep, c) : w_amino_occ(ix, ep, c);
  sp = ix.c_lo[c] + ra;
  ep = ix.c_lo[c] + rb - 1;
}
// backstep (fm_index.rs:585-593): LF(row), or 0 when the row holds '$'
template <int ALPHA>
__device__ __forceinline__ uint64_t w_lf_backstep(const WideView& ix, uint64_t row) {
  if (ALPHA == 0) {
    const uint32_t c = w_dna_code_at(ix, row);
    if (c >= uint32_t(DNA_SENTINEL)) return 0;
    return ix.c_lo[c] + w_dna_occ(ix, row, c) - 1;
  }
  const uint32_t s = w_amino_symbol_at(ix, row);
  if (s == uint32_t(AMINO_SENTINEL) || s > 21) return 0;
  return ix.c_lo[s] + w_amino_occ(ix, row, s) - 1;
}
__device__ __forceinline__ bool w_row_is_sampled(const WideView& ix, uint64_t row) {
  return ix.sa_pow2 ? (row & (ix.sa_ratio - 1)) == 0 : (row % ix.sa_ratio) == 0;
}
// CompressedSuffixArray::reconstruct_value (compressed_suffix_array.rs:76-106)
__device__ __forceinline__ uint64_t w_sa_sample(const WideView& ix, uint64_t row) {
  const uint64_t e = ix.sa_pow2 ? (row >> ix.sa_ratio_shift) : (row / ix.sa_ratio);
  const uint64_t bit = e * ix.sa_bits;  // < 2^64 for every bwt_len < 2^57
  const uint64_t w = bit >> 6;
  const uint32_t s = uint32_t(bit & 63);
  AWRY_CHK(w + (s + ix.sa_bits > 64 ? 1 : 0) < ix.n_sa_words);
  uint64_t v = __ldg(ix.sa_words + w) >> s;
  if (s + ix.sa_bits > 64) v |= __ldg(ix.sa_words + w + 1) << (64 - s);
  return ix.sa_bits >= 64 ? v : (v & ((1ull << ix.sa_bits) - 1));
}
__device__ __forceinline__ void w_map_location(const WideView& ix, uint64_t loc, uint64_t* out2) {  // sequence_index.rs:108-141 (Q4)
  uint32_t lo = 0;
  if (ix.n_seqs > 1) {
    uint32_t hi = ix.n_seqs - 1;
    while (lo < hi) {
      const uint32_t mid = (lo + hi + 1) >> 1;
      if (__ldg(ix.seq_starts + mid) <= loc)
        lo = mid;
      else
        hi = mid - 1;
    }
  }
  out2[0] = lo;
  out2[1] = loc - __ldg(ix.seq_starts + lo);
}

// ------------------------------------------------------------------ seed table

__device__ __forceinline__ uint32_t w_amino_digit_to_sym(uint32_t d) { return d < 19 ? d + 1 : 21; }

template <int ALPHA>
__global__ void w_build_table_kernel(WideView ix, ulonglong2* __restrict__ table, uint64_t n, uint32_t k) {
  constexpr uint32_t B = ALPHA == 0 ? 4 : 20;
  const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  uint64_t rest = idx;
  uint32_t d = uint32_t(rest % B);
  rest /= B;
  const uint32_t c = ALPHA == 0 ? d : w_amino_digit_to_sym(d);
  uint64_t sp = ix.c_lo[c], ep = ix.c_hi[c];
  for (uint32_t j = 1; j < k && sp <= ep; j++) {
    d = uint32_t(rest % B);
    rest /= B;
    w_lf_update<ALPHA>(ix, sp, ep, ALPHA == 0 ? d : w_amino_digit_to_sym(d));
  }
  table[idx] = sp <= ep ? make_ulonglong2(sp, ep) : make_ulonglong2(1, 0);
}

cudaError_t launch_build_table_wide(const WideView& ix, ulonglong2* d_table, uint32_t k, cudaStream_t s) {
  const uint64_t n = table_entries(int(ix.alphabet), k);
  const unsigned grid = unsigned((n + 255) / 256);
  if (ix.alphabet == 0)
    w_build_table_kernel<0><<<grid, 256, 0, s>>>(ix, d_table, n, k);
  else
    w_build_table_kernel<1><<<grid, 256, 0, s>>>(ix, d_table, n, k);
  COUNT_LAUNCH();
  return cudaGetLastError();
}

// ------------------------------------------------------------------ backward search, one thread per query

template <int MODE>
__device__ __forceinline__ void w_store_result(void* out, uint64_t q, uint64_t sp, uint64_t ep) {
  const bool empty = sp > ep;
  if (MODE == OUT_COUNT_U64)
    reinterpret_cast<uint64_t*>(out)[q] = empty ? 0ull : ep - sp + 1ull;  // search.rs:66-71
  else if (MODE == OUT_RANGE_U64)
    reinterpret_cast<ulonglong2*>(out)[q] = empty ? make_ulonglong2(1, 0) : make_ulonglong2(sp, ep);
  else  // (sp, count): 16 bytes per query in wide mode
    reinterpret_cast<ulonglong2*>(out)[q] = empty ? make_ulonglong2(1, 0) : make_ulonglong2(sp, ep - sp + 1ull);
}

template <int ALPHA, int MODE>
__global__ void __launch_bounds__(256) w_search_kernel(WideView ix, const uint64_t* __restrict__ qwords,
                                                       const uint64_t* __restrict__ qoff, uint64_t nq,
                                                       void* __restrict__ out, ByteRange br) {
  constexpr uint32_t SENT = ALPHA == 0 ? uint32_t(DNA_SENTINEL) : uint32_t(AMINO_SENTINEL);
  const uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t q = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; q < nq; q += stride) {
    const uint64_t o0 = qoff[q];
    const uint32_t len = checked_len(o0, qoff[q + 1], br);
    uint64_t sp = 1, ep = 0;
    if (len != 0) {
      QueryStream<ALPHA> qs;
      qs.open(qwords, uint32_t(q), o0);
      uint32_t left = 0;
      // seed from the k-mer table when the last k symbols are all encoding symbols (kmer_lookup_table.rs:90-110
      // recomputes these k-1 steps), else the single-symbol range (search.rs:43-48)
      const uint32_t k = ix.kmer_len;
      bool seeded = false;
      if (k != 0 && len >= k) {
        uint64_t idx = 0;
        bool ok = true;
        const uint64_t w = qs.w;
        if (ALPHA == 0) {
          ok = (w & (0xCCCCCCCCCCCCCCCCull >> (4 * (16 - k)))) == 0;
          idx = kmer_index(w, k);
        } else {
          uint64_t mult = 1;
          for (uint32_t j = 0; j < k; j++) {
            const uint32_t s = uint32_t(w >> (8 * j)) & 0xff;
            ok &= (s != uint32_t(AMINO_X)) & (s != uint32_t(AMINO_SENTINEL));
            idx += uint64_t(s == 21 ? 19 : s - 1) * mult;
            mult *= 20;
          }
        }
        if (ok) {
          AWRY_CHK(idx < ix.n_table);
          const ulonglong2 r = ix.table[idx];
          sp = r.x;
          ep = r.y;
          for (uint32_t j = 0; j < k; j++) qs.next(qwords);
          left = len - k;
          seeded = true;
        }
      }
      if (!seeded) {
        const uint32_t c = qs.next(qwords);
        if (c != SENT) {
          sp = ix.c_lo[c];
          ep = ix.c_hi[c];
          left = len - 1;
        }
      }
      while (left != 0 && sp <= ep) {  // early break: fm_index.rs:409-416 / :425-433
        const uint32_t c = qs.next(qwords);
        left--;
        if (c == SENT || (ALPHA == 1 && c > 21) || (ALPHA == 0 && c > 4)) {
          sp = 1;
          ep = 0;
          break;
        }
        w_lf_update<ALPHA>(ix, sp, ep, c);
      }
    }
    w_store_result<MODE>(out, q, sp, ep);
  }
}

// Nucleotide, cooperative: 4 lanes per query on the 64-B block (one LDG.128 each = ONE line request per LF
// step instead of the eight of the one-thread kernel above), both ranks of a step from one block load when they
// share the block, the u32 partial ranks reduced with xor-shuffles and the superblock count added after the
// reduction; queries handed out dynamically per lane group (tickets), groups refill in-loop, the warp leaves by
// vote.  The shape of search_dna_kernel<4> / search_amino_kernel on 64-bit rows; ambiguity symbols take the
// scalar step.  (No pair index for wide indexes yet: one symbol per block read.)
template <int MODE>
__global__ void __launch_bounds__(256, 6)
    w_search_dna_coop_kernel(WideView ix, const uint64_t* __restrict__ qwords, const uint64_t* __restrict__ qoff,
                             uint64_t nq, void* __restrict__ out, uint32_t* __restrict__ ticket, uint32_t ticket_sz,
                             ByteRange br) {
  constexpr uint32_t NONE = 0xffffffffu, FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31, sub = lane & 3, gbase = lane - sub;
  const uint32_t gmask = 0xfu << gbase;
  const uint32_t nq32 = uint32_t(nq);
  uint32_t q = 0, q_end = 0;
  bool more = true;
  uint32_t cur = NONE, left = 0;
  uint64_t sp = 1, ep = 0;
  QueryStream<0> qs;
  qs.w = qs.wnext = 0;
  qs.widx = 0;
  qs.inword = 0;
  for (;;) {
    if (left == 0 || sp > ep) {
      if (cur != NONE && sub == 0) w_store_result<MODE>(out, cur, sp, ep);
      cur = NONE;
      left = 0;
      if (q == q_end && more) {
        uint32_t t = 0;
        if (sub == 0) t = atomicAdd(ticket, ticket_sz);
        t = __shfl_sync(gmask, t, gbase);
        more = t < nq32;
        q = more ? t : 0u;
        q_end = more ? (nq32 - t < ticket_sz ? nq32 : t + ticket_sz) : 0u;
      }
      if (q < q_end) {
        cur = q++;
        const uint64_t o0 = qoff[cur];
        const uint32_t len = checked_len(o0, qoff[cur + 1], br);
        sp = 1;
        ep = 0;
        if (len != 0) {
          qs.open(qwords, cur, o0);
          const uint32_t k = ix.kmer_len;
          const uint64_t w = qs.w;
          if (k != 0 && len >= k && (w & (0xCCCCCCCCCCCCCCCCull >> (4 * (16 - k)))) == 0) {
            uint64_t idx = 0;
            idx = kmer_index(w, k);
            AWRY_CHK(idx < ix.n_table);
            const ulonglong2 r = ix.table[idx];
            sp = r.x;
            ep = r.y;
#pragma unroll 1
            for (uint32_t j = 0; j < k; j++) qs.next(qwords);
            left = len - k;
          } else {
            const uint32_t c = qs.next(qwords);
            if (c <= uint32_t(DNA_N)) {
              sp = ix.c_lo[c];
              ep = ix.c_hi[c];
              left = len - 1;
            }
          }
        }
      }
    }
    if (__all_sync(FULL, cur == NONE && q == q_end && !more)) break;
    const bool active = left != 0 && sp <= ep;
    uint32_t c = 7;
    if (active) {
      c = qs.next(qwords);
      left--;
    }
    uint32_t ra = 0, rb = 0;
    const uint64_t pa = sp - 1, pb = ep;
    const uint64_t ba = pa >> 7, bb = pb >> 7;
    if (active && c < 4) {
      AWRY_CHK(ba * DNA_BLOCK_UINT4 + 3 < ix.n_blocks_u4 && bb * DNA_BLOCK_UINT4 + 3 < ix.n_blocks_u4);
      LaneChunks<4> x;
      x.load(ix.blocks + size_t(ba) * DNA_BLOCK_UINT4, sub);
      const uint32_t m0 = (c & 1) ? ~0u : 0u, m1 = (c & 2) ? ~0u : 0u;
      ra = dna_partial_rank<4>(x, sub, uint32_t(pa) & 127, c, m0, m1);
      if (bb != ba) x.load(ix.blocks + size_t(bb) * DNA_BLOCK_UINT4, sub);
      rb = dna_partial_rank<4>(x, sub, uint32_t(pb) & 127, c, m0, m1);
    }
    ra += __shfl_xor_sync(FULL, ra, 1);
    rb += __shfl_xor_sync(FULL, rb, 1);
    ra += __shfl_xor_sync(FULL, ra, 2);
    rb += __shfl_xor_sync(FULL, rb, 2);
    if (active) {
      if (c < 4) {  // block counts are relative to the block's superblock
        const uint64_t sa = __ldg(ix.sb_counts + (ba >> (ix.sb_shift - 7)) * SB_STRIDE + c);
        const uint64_t sb = __ldg(ix.sb_counts + (bb >> (ix.sb_shift - 7)) * SB_STRIDE + c);
        sp = ix.c_lo[c] + sa + ra;
        ep = ix.c_lo[c] + sb + rb - 1;
      } else if (c == uint32_t(DNA_N)) {
        w_lf_update<0>(ix, sp, ep, c);  // rare: every lane of the group runs the scalar step
      } else {
        sp = 1;  // sentinel in a query: refused by the prepass; stay defined
        ep = 0;
      }
    }
  }
}

template <int MODE>
static cudaError_t launch_search_wide_coop(const WideView& ix, const uint64_t* d_qwords, const uint64_t* d_qoff, uint64_t nq,
                                           void* d_out, uint32_t* d_ticket, uint32_t avg_len, ByteRange br, int sm_count,
                                           cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, 4, s);
  if (e != cudaSuccess) return e;
  auto kern = w_search_dna_coop_kernel<MODE>;
  int per_sm = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0)) != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  const uint64_t need_blocks = (nq * 4 + 255) / 256;
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(sm_count) * per_sm, need_blocks)));
  kern<<<grid, 256, 0, s>>>(ix, d_qwords, d_qoff, nq, d_out, d_ticket, ticket_size(avg_len), br);
  COUNT_LAUNCH();
  return cudaGetLastError();
}

template <int ALPHA>
static cudaError_t launch_search_wide_a(const WideView& ix, const uint64_t* d_qwords, const uint64_t* d_qoff, uint64_t nq,
                                        SearchOut mode, void* d_out, ByteRange br, int sm_count, cudaStream_t s) {
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(sm_count) * 8, (nq + 255) / 256)));
  switch (mode) {
    case OUT_COUNT_U64: w_search_kernel<ALPHA, OUT_COUNT_U64><<<grid, 256, 0, s>>>(ix, d_qwords, d_qoff, nq, d_out, br); break;
    case OUT_RANGE_U64: w_search_kernel<ALPHA, OUT_RANGE_U64><<<grid, 256, 0, s>>>(ix, d_qwords, d_qoff, nq, d_out, br); break;
    default: w_search_kernel<ALPHA, OUT_SP_CNT_U32><<<grid, 256, 0, s>>>(ix, d_qwords, d_qoff, nq, d_out, br); break;
  }
  COUNT_LAUNCH();
  return cudaGetLastError();
}

cudaError_t launch_search_wide(const WideView& ix, const uint64_t* d_qwords, const uint64_t* d_qoff, uint64_t nq,
                               SearchOut mode, void* d_out, uint32_t* d_scratch, const SearchVariant& v, int sm_count,
                               cudaStream_t s) {
  if (nq == 0) return cudaSuccess;
  if (nq >= (1ull << 32) - (1u << 24)) return cudaErrorInvalidValue;  // 32-bit packed-word indices and ticket counter
  const ByteRange br{v.b_lo, v.b_hi};
  if (ix.alphabet == 0 && v.lanes != -1) {  // lanes == -1: the one-thread kernels (cross-check)
    switch (mode) {
      case OUT_COUNT_U64: return launch_search_wide_coop<OUT_COUNT_U64>(ix, d_qwords, d_qoff, nq, d_out, d_scratch, v.avg_len, br, sm_count, s);
      case OUT_RANGE_U64: return launch_search_wide_coop<OUT_RANGE_U64>(ix, d_qwords, d_qoff, nq, d_out, d_scratch, v.avg_len, br, sm_count, s);
      default: return launch_search_wide_coop<OUT_SP_CNT_U32>(ix, d_qwords, d_qoff, nq, d_out, d_scratch, v.avg_len, br, sm_count, s);
    }
  }
  return ix.alphabet == 0 ? launch_search_wide_a<0>(ix, d_qwords, d_qoff, nq, mode, d_out, br, sm_count, s)
                          : launch_search_wide_a<1>(ix, d_qwords, d_qoff, nq, mode, d_out, br, sm_count, s);
}

// ------------------------------------------------------------------ locate

struct WideCountOf {
  const ulonglong2* r;
  uint64_t nq;
  __host__ __device__ uint64_t operator()(uint64_t i) const { return i < nq ? r[i].y : 0ull; }
};

cudaError_t scan_hit_offsets_wide(const void* d_sp_cnt, uint64_t nq, uint64_t* d_hit_off, void* d_temp, size_t& temp_bytes,
                                  cudaStream_t s) {
  thrust::counting_iterator<uint64_t> idx(0);
  auto in = thrust::make_transform_iterator(idx, WideCountOf{static_cast<const ulonglong2*>(d_sp_cnt), nq});
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, in, d_hit_off, nq + 1, s);
  if (d_temp != nullptr) COUNT_LAUNCH();
  return e;
}

// CSR -> one BWT row per hit in the first word of each output slot (a lane writes the first 8 rows of its
// query, longer intervals are written by the whole warp; see expand_rows_kernel)
__global__ void __launch_bounds__(256)
    w_expand_rows_kernel(const ulonglong2* __restrict__ sp_cnt, const uint64_t* __restrict__ hit_off, uint64_t nq,
                         uint64_t* __restrict__ out, uint32_t slot) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) >> 5;
  const uint64_t nwarps = (gridDim.x * uint64_t(blockDim.x)) >> 5;
  for (uint64_t base = warp * 32; base < nq; base += nwarps * 32) {
    const uint64_t q = base + lane;
    uint64_t sp = 0, cnt = 0, off = 0;
    if (q < nq) {
      const ulonglong2 r = sp_cnt[q];
      sp = r.x;
      cnt = r.y;
      off = hit_off[q];
    }
    const uint64_t small = cnt < 8 ? cnt : 8;
    for (uint64_t i = 0; i < small; i++) out[(off + i) * slot] = sp + i;
    uint32_t big = __ballot_sync(0xffffffffu, cnt > 8);
    while (big) {
      const int L = __ffs(big) - 1;
      big &= big - 1;
      const uint64_t s = __shfl_sync(0xffffffffu, sp, L), c = __shfl_sync(0xffffffffu, cnt, L),
                     o = __shfl_sync(0xffffffffu, off, L);
      for (uint64_t i = 8 + lane; i < c; i += 32) out[(o + i) * slot] = s + i;
    }
  }
}

// locate_string's inner loop (fm_index.rs:521-537), one thread per hit: LF-walk to a sampled row, then
// (sample + steps) % bwt_len -- the wrap is the walk through the '$' row (fm_index.rs:587-589)
template <int ALPHA, bool MAP>
__global__ void __launch_bounds__(256) w_walk_kernel(WideView ix, uint64_t n_hits, uint64_t* __restrict__ out) {
  constexpr int SLOT = MAP ? 2 : 1;
  const uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t h = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; h < n_hits; h += stride) {
    uint64_t row = out[SLOT * h], steps = 0;
    while (!w_row_is_sampled(ix, row)) {
      row = w_lf_backstep<ALPHA>(ix, row);
      steps++;
    }
    const uint64_t loc = (w_sa_sample(ix, row) + steps) % ix.bwt_len;
    if (MAP)
      w_map_location(ix, loc, out + SLOT * h);
    else
      out[h] = loc;
  }
}

cudaError_t launch_walk_wide(const WideView& ix, const void* d_sp_cnt, const uint64_t* d_hit_off, uint64_t nq, uint64_t n_hits,
                             uint64_t* d_hits_pairs, uint64_t* d_locs, int sm_count, cudaStream_t s) {
  if (n_hits == 0 || nq == 0) return cudaSuccess;
  const bool map = d_hits_pairs != nullptr;
  uint64_t* out = map ? d_hits_pairs : d_locs;
  unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(sm_count) * 8, (nq + 255) / 256)));
  w_expand_rows_kernel<<<grid, 256, 0, s>>>(static_cast<const ulonglong2*>(d_sp_cnt), d_hit_off, nq, out, map ? 2u : 1u);
  COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>(uint64_t(sm_count) * 16, (n_hits + 255) / 256)));
  if (ix.alphabet == 0) {
    if (map)
      w_walk_kernel<0, true><<<grid, 256, 0, s>>>(ix, n_hits, out);
    else
      w_walk_kernel<0, false><<<grid, 256, 0, s>>>(ix, n_hits, out);
  } else {
    if (map)
      w_walk_kernel<1, true><<<grid, 256, 0, s>>>(ix, n_hits, out);
    else
      w_walk_kernel<1, false><<<grid, 256, 0, s>>>(ix, n_hits, out);
  }
  COUNT_LAUNCH();
  return cudaGetLastError();
}

// ------------------------------------------------------------------ single steps

template <int ALPHA>
__global__ void w_single_update_kernel(WideView ix, uint64_t sp, uint64_t ep, uint32_t c, uint64_t* out) {
  w_lf_update<ALPHA>(ix, sp, ep, c);
  out[0] = sp;
  out[1] = ep;
}
template <int ALPHA>
__global__ void w_single_backstep_kernel(WideView ix, uint64_t row, uint64_t* out) {
  out[0] = w_lf_backstep<ALPHA>(ix, row);
}
cudaError_t launch_single_update_wide(const WideView& ix, uint64_t sp, uint64_t ep, uint32_t dsym, uint64_t* d_out2,
                                      cudaStream_t s) {
  if (ix.alphabet == 0)
    w_single_update_kernel<0><<<1, 1, 0, s>>>(ix, sp, ep, dsym, d_out2);
  else
    w_single_update_kernel<1><<<1, 1, 0, s>>>(ix, sp, ep, dsym, d_out2);
  COUNT_LAUNCH();
  return cudaGetLastError();
}
cudaError_t launch_single_backstep_wide(const WideView& ix, uint64_t row, uint64_t* d_out, cudaStream_t s) {
  if (ix.alphabet == 0)
    w_single_backstep_kernel<0><<<1, 1, 0, s>>>(ix, row, d_out);
  else
    w_single_backstep_kernel<1><<<1, 1, 0, s>>>(ix, row, d_out);
  COUNT_LAUNCH();
  return cudaGetLastError();
}

// The reference's own (incomplete) k-mer table for FmIndex::save (see ref_table_kernel in kernels.cu)
template <int ALPHA>
__global__ void w_ref_table_kernel(WideView ix, uint64_t first, uint64_t count, uint32_t k, ulonglong2* __restrict__ out) {
  constexpr uint32_t ENC = ALPHA == 0 ? 4 : 20;
  const uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (t >= count) return;
  uint64_t rest = first + t;
  uint32_t d = uint32_t(rest % ENC);
  rest /= ENC;
  bool visited = d != 0;
  uint64_t sp = 1, ep = 0;
  if (visited) {
    const uint32_t c = ALPHA == 0 ? d - 1 : d;
    sp = ix.c_lo[c];
    ep = ix.c_hi[c];
  }
  for (uint32_t j = 1; j < k && visited; j++) {
    d = uint32_t(rest % ENC);
    rest /= ENC;
    if (d == 0) {
      visited = false;
      break;
    }
    w_lf_update<ALPHA>(ix, sp, ep, ALPHA == 0 ? d - 1 : d);
  }
  out[t] = visited ? make_ulonglong2(sp, ep) : make_ulonglong2(1, 0);
}
cudaError_t launch_ref_table_wide(const WideView& ix, uint64_t first, uint64_t count, uint32_t k, void* d_out, cudaStream_t s) {
  if (count == 0) return cudaSuccess;
  const unsigned grid = unsigned((count + 255) / 256);
  if (ix.alphabet == 0)
    w_ref_table_kernel<0><<<grid, 256, 0, s>>>(ix, first, count, k, static_cast<ulonglong2*>(d_out));
  else
    w_ref_table_kernel<1><<<grid, 256, 0, s>>>(ix, first, count, k, static_cast<ulonglong2*>(d_out));
  COUNT_LAUNCH();
  return cudaGetLastError();
}

}  // namespace awry
