// layout.cuh -- device-side index layout and rank primitives (sm_100a).
//
// Replaces the reference's rank structure: the 256-row windowed bit-vector blocks of
// /root/reference/src/bwt.rs:12-25 (160 B nucleotide / 352 B amino, u64 milestones) and the
// AVX2/NEON predicate + masked popcount of simd_instructions.rs:78-121.
//
// Device layout.  Row pointers are 32-bit while bwt_len < 2^32 - 256 (every BASELINE config); larger indexes
// use the SAME blocks through WideView below (64-bit row pointers, block counts relative to 2^31-row
// superblocks) and the kernels of kernels_wide.cu.
//
//   NUCLEOTIDE  block = 128 BWT rows = 64 B = 4 chunks of 16 B (one uint4 / one LDG.128 each)
//               chunk j = { p0, p1, p2, cnt_j }   p_b = bit-plane b of rows 32j..32j+31
//               (bit t = row 32j+t), cnt_j = #symbol j in BWT[0 .. block start)  (j: A,C,G,T).
//               Row code (3 planes) = device symbol: A0 C1 G2 T3 N4 $5, padding rows 7.
//               The N milestone is derived: start - (A+C+G+T) - [$ row < start].
//               => 4 bits per BWT row all-in; one LF step reads one 64-B aligned block.
//
//   NUCLEOTIDE PAIR INDEX (accelerator, derived from the blocks above at load time)
//               block = 96 BWT rows = 128 B = 4 lane slices of 32 B (one LDG.256 each)
//               row code (5 planes) = 4*a + b with a = BWT[row], b = BWT[LF(row)], both in
//               A,C,G,T; plane 4 marks rows where either symbol is N or '$' (they match no pair).
//               slice t < 3 = { p0..p4 of rows 32t..32t+31, cnt[3t], cnt[3t+1], cnt[3t+2] }
//               slice 3     = { cnt[9] .. cnt[15], spare }
//               cnt[p] = #rows with pair code p in BWT[0 .. block start).
//               One access advances the backward search by TWO query symbols:
//                 sp'' = C2[a,b] + Occ2(ab, sp-1),  ep'' = C2[a,b] + Occ2(ab, ep) - 1,
//                 C2[a,b] = C[b] + Occ(b, C[a]-1)
//               (two applications of fm_index.rs:559-582 composed).
//
//   NUCLEOTIDE WALK BLOCKS + POSITION-SAMPLED SA (memory-lean locate; derived at load time)
//               The file samples the suffix array by ROW (row % ratio == 0, compressed_suffix_array.rs:109-111),
//               so a walk is geometric and unbounded.  Sampling by TEXT POSITION (SA[row] % r == 0) bounds
//               it by r - 1 steps; it needs a mark bit per row and a rank over the marks.  Both are DERIVED at
//               load time, so r is the library's choice, not the file's: r = min(file ratio, 4) by default.
//               block = 224 BWT rows = 128 B = 4 lane slices of 32 B (one LDG.256 each)
//               words 4g .. 4g+3 (g = 0..6) = { p0, p1, p2, mark } of rows 32g .. 32g+31
//               (p_b = bit-plane b of the device row code, mark bit t = "SA[row 32g+t] % ratio == 0")
//               words 28 .. 31 = #A, #C, #G, #T in BWT[0 .. block start)
//               slice t < 3 holds row groups 2t and 2t+1, slice 3 holds group 6 and the counts: the BWT symbol,
//               its rank AND the stop test of an LF step come from ONE aligned 128-B line.
//               walk_rank[blk] = marked rows before the block; pos_samples[i] = SA[i-th marked row] / r.
//               4.57 + 32/r bits per row (12.6 at r = 4: 4.9 GB at 3.1 G rows) against 32 for the unsampled array.
//
//   NUCLEOTIDE TEXT, REVERSED (count accelerator; derived at load time from the blocks + the unsampled array)
//               rtext[i] = device symbol of text[n - 1 - i], 4 bits each (low nibble = even i): the format of a
//               packed query, whose symbol 0 is the query's LAST one -- so "the rest of the query against the
//               text in front of position p" is a word-wise XOR of two forward streams.  N and '$' keep their
//               codes (4, 5): an A/C/G/T query symbol never equals them, as in the backward search.
//               n / 2 bytes (1.55 GB at 3.1 G symbols).
//
//   AMINO       block = 64 rows = 128 B = 4 lane slices of 32 B (one LDG.256 each; one line per step)
//               slice t < 2 = { p0..p4 of rows 32t..32t+31, cnt[3t], cnt[3t+1], cnt[3t+2] }
//               slice 2     = { cnt[6] .. cnt[13] },  slice 3 = { cnt[14] .. cnt[21] }
//               cnt[s-1] = #symbol s in BWT[0 .. block start), s = reference symbol index 1..21.
//               Row code (5 planes) = reference symbol index (0 = '$' and padding rows).
//
// Ranks are INCLUSIVE, Occ(c,i) = #c in BWT[0..=i], as in bwt.rs:114-135 / fm_index.rs:559-582.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// ---- checked build (libawry_b200_checked.so, `make checked`): compute-sanitizer is closed on the GPU pool this
// was developed on, so every load whose address is computed from index or query DATA -- rank blocks, pair
// blocks, seed table, suffix arrays, walk blocks, packed query words -- is bounds-checked by the code itself
// against the array sizes the views carry; a violation is a device assert (the launch fails with
// cudaErrorAssert, file and line on stderr).  tests/test_gpu_checked.py runs the GPU suite against that build.
#ifdef AWRY_B200_CHECKED
#include <cassert>
#define AWRY_CHK(cond) assert(cond)
#else
#define AWRY_CHK(cond) ((void)0)
#endif

namespace awry {

// ---- device symbol numbering -------------------------------------------------------------
// nucleotide: A0 C1 G2 T3 N4 $5 (reference indices, alphabet.rs:228-235: $0 A1 C2 G3 N4 T5)
// amino:      reference index unchanged (alphabet.rs:174-196), X = 20 is the ambiguity symbol
constexpr int DNA_N = 4, DNA_SENTINEL = 5, DNA_PAD = 7;
constexpr int AMINO_X = 20, AMINO_SENTINEL = 0;

constexpr uint32_t DNA_ROWS_PER_BLOCK = 128, DNA_BLOCK_UINT4 = 4;
constexpr uint32_t AMINO_ROWS_PER_BLOCK = 64, AMINO_BLOCK_UINT4 = 8;
constexpr uint32_t PAIR_ROWS_PER_BLOCK = 96, PAIR_BLOCK_UINT4 = 8;
constexpr uint32_t WALK_ROWS_PER_BLOCK = 224, WALK_BLOCK_UINT4 = 8;

struct WideView;

struct IndexView {
  const uint4* __restrict__ blocks;
  const uint64_t* __restrict__ sa_words;
  const uint2* __restrict__ table;         // k-mer seeds: (sp, ep), empty = (1,0)
  const uint64_t* __restrict__ seq_starts;
  const uint4* __restrict__ pair_blocks;   // nucleotide two-step accelerator, or nullptr
  const uint32_t* __restrict__ full_sa;    // unsampled suffix array (locate accelerator), or nullptr
  const uint4* __restrict__ walk_blocks;   // nucleotide walk blocks (planes + position marks), or nullptr
  const uint32_t* __restrict__ walk_rank;  // marked rows before each walk block
  const uint32_t* __restrict__ pos_samples;  // SA[marked row] / lean_ratio, in row order
  const uint8_t* __restrict__ rtext;       // the indexed text REVERSED, one device symbol per 4 bits (the packed
                                           // query format): rtext[i] = text[bwt_len - 1 - i], i even = low nibble;
                                           // derived from BWT + unsampled SA at load time, or nullptr
  uint32_t lean_ratio;                     // a row is marked iff SA[row] % lean_ratio == 0 (the library's choice:
                                           // the array is derived, not stored -- see finish_replica0)
  uint32_t c2[16];                         // C2[4a+b]
  uint32_t c_lo[24];                       // C[c]      by device symbol (search.rs:43-48)
  uint32_t c_hi[24];                       // C[c+1]-1  by device symbol
  uint32_t bwt_len;
  uint32_t dollar_row;                     // row whose BWT symbol is '$' (SA = 0)
  uint32_t sa_ratio;
  uint32_t sa_pow2;                        // 1 if sa_ratio is a power of two
  uint32_t sa_ratio_shift;                 // log2(sa_ratio) when sa_pow2
  uint32_t sa_bits;
  uint32_t kmer_len;                       // 0 = no seed table
  uint32_t n_seqs;
  uint32_t alphabet;
  const WideView* wide;                    // HOST pointer, never read on the device: set for indexes with
                                           // bwt_len >= 2^32 - 256; the launchers then run kernels_wide.cu
  // array sizes in elements (what AWRY_CHK checks against in the checked build)
  uint64_t n_blocks_u4, n_pair_u4, n_table, n_sa_words, n_full_sa, n_walk_u4, n_walk_rank, n_pos_samples;
  uint64_t n_rtext;  // bytes
};

// ---- wide indexes: SearchPtr = u64 (search.rs:7), suffix-array elements up to 64 bits
// (compressed_suffix_array.rs:124-130).  The block layouts above are unchanged; their u32 block-start counts
// are relative to the start of the block's SUPERBLOCK (2^31 rows), whose absolute counts live in sb_counts.
constexpr uint32_t SB_SHIFT = 31;   // rows per superblock = 2^31 (WideView::sb_shift; AWRY_B200_SB_SHIFT lowers it so that
                                    // tests cross superblock borders on small indexes)
constexpr uint32_t SB_STRIDE = 24;  // u64 counts per superblock: nucleotide A,C,G,T at 0..3; amino symbol index s at s
struct WideView {
  const uint4* __restrict__ blocks;
  const uint64_t* __restrict__ sa_words;
  const ulonglong2* __restrict__ table;      // k-mer seeds (sp, ep), empty = (1, 0)
  const uint64_t* __restrict__ seq_starts;
  const uint64_t* __restrict__ sb_counts;    // [superblock][SB_STRIDE]
  uint64_t c_lo[24];                         // C[c]      by device symbol
  uint64_t c_hi[24];                         // C[c+1]-1
  uint64_t bwt_len;
  uint64_t dollar_row;
  uint32_t sa_ratio, sa_pow2, sa_ratio_shift, sa_bits, kmer_len, n_seqs, alphabet;
  uint32_t sb_shift;                         // log2(rows per superblock), >= 8
  uint64_t n_blocks_u4, n_table, n_sa_words, n_sb;  // array sizes in elements (checked build)
};

// 128-/256-bit read-only loads that do not allocate in L1: every block is touched once per
// step by one query, so L1 residency buys nothing and only evicts the query words.
__device__ __forceinline__ uint4 ldg128(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
struct u32x8 {
  uint32_t v[8];
};
__device__ __forceinline__ u32x8 ldg256(const void* p) {  // sm_100+: LDG.E.256
  u32x8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]),
                 "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}

// mask of bits 0..=local for the 32-row chunk j (masked_popcount, simd_instructions.rs:96-121)
__device__ __forceinline__ uint32_t chunk_mask(uint32_t local, uint32_t j) {
  int n = int(local) + 1 - int(32 * j);
  n = n < 0 ? 0 : n;
  return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
}

// ---- nucleotide --------------------------------------------------------------------------

// rows of a chunk whose 3-bit code equals c; (m0,m1,m2) = code bits of c spread to full words
__device__ __forceinline__ uint32_t dna_match(const uint4& ch, uint32_t m0, uint32_t m1, uint32_t m2) {
  return ~(ch.x ^ m0) & ~(ch.y ^ m1) & ~(ch.z ^ m2);
}

struct DnaBlockRegs {
  uint4 ch[4];
};
__device__ __forceinline__ DnaBlockRegs dna_load_block(const uint4* __restrict__ blocks, uint64_t blk) {
  DnaBlockRegs b;
  const uint4* p = blocks + size_t(blk) * DNA_BLOCK_UINT4;
#pragma unroll
  for (int j = 0; j < 4; j++) b.ch[j] = ldg128(p + j);
  return b;
}
__device__ __forceinline__ DnaBlockRegs dna_load_block(const IndexView& ix, uint32_t blk) {
  AWRY_CHK(uint64_t(blk) * DNA_BLOCK_UINT4 + 3 < ix.n_blocks_u4);
  return dna_load_block(ix.blocks, blk);
}
// milestone of device symbol c (0..4) at the start of block blk
__device__ __forceinline__ uint32_t dna_milestone(const IndexView& ix, const DnaBlockRegs& b,
                                                  uint32_t blk, uint32_t c) {
  if (c < 4) return c == 0 ? b.ch[0].w : c == 1 ? b.ch[1].w : c == 2 ? b.ch[2].w : b.ch[3].w;
  uint32_t start = blk * DNA_ROWS_PER_BLOCK;
  return start - (b.ch[0].w + b.ch[1].w + b.ch[2].w + b.ch[3].w) - (ix.dollar_row < start ? 1u : 0u);
}
// Occ(c, pos) within an already loaded block (c in 0..4)
__device__ __forceinline__ uint32_t dna_occ_in_block(const IndexView& ix, const DnaBlockRegs& b,
                                                     uint32_t blk, uint32_t local, uint32_t c) {
  uint32_t m0 = (c & 1) ? ~0u : 0u, m1 = (c & 2) ? ~0u : 0u, m2 = (c & 4) ? ~0u : 0u;
  uint32_t r = dna_milestone(ix, b, blk, c);
#pragma unroll
  for (int j = 0; j < 4; j++) r += __popc(dna_match(b.ch[j], m0, m1, m2) & chunk_mask(local, j));
  return r;
}
__device__ __forceinline__ uint32_t dna_code_at(const DnaBlockRegs& b, uint32_t local) {
  uint32_t j = local >> 5, t = local & 31;
  uint4 ch = j == 0 ? b.ch[0] : j == 1 ? b.ch[1] : j == 2 ? b.ch[2] : b.ch[3];
  return ((ch.x >> t) & 1u) | (((ch.y >> t) & 1u) << 1) | (((ch.z >> t) & 1u) << 2);
}

// ---- amino -------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t amino_match(const u32x8& ch, uint32_t s) {
  uint32_t r = ~0u;
#pragma unroll
  for (int p = 0; p < 5; p++) r &= ((s >> p) & 1u) ? ch.v[p] : ~ch.v[p];
  return r;
}
// 32-bit word of an amino block that holds the block-start count of symbol s (1..21)
__host__ __device__ __forceinline__ uint32_t amino_count_word(uint32_t s) {
  uint32_t slot = s - 1;
  return slot < 3 ? 5 + slot : slot < 6 ? 8 + 5 + (slot - 3) : slot < 14 ? 16 + (slot - 6) : 24 + (slot - 14);
}
__device__ __forceinline__ uint32_t amino_milestone(const IndexView& ix, uint32_t blk, uint32_t s) {
  AWRY_CHK(uint64_t(blk) * AMINO_BLOCK_UINT4 + 7 < ix.n_blocks_u4 && s >= 1 && s <= 21);
  const uint32_t* words =
      reinterpret_cast<const uint32_t*>(ix.blocks + size_t(blk) * AMINO_BLOCK_UINT4);
  return __ldg(words + amino_count_word(s));
}
// Occ(s, pos) for reference symbol index s in 1..21
__device__ __forceinline__ uint32_t amino_occ(const IndexView& ix, uint32_t pos, uint32_t s) {
  uint32_t blk = pos >> 6, local = pos & 63;
  const char* base = reinterpret_cast<const char*>(ix.blocks + size_t(blk) * AMINO_BLOCK_UINT4);
  uint32_t r = amino_milestone(ix, blk, s);
  u32x8 c0 = ldg256(base);
  r += __popc(amino_match(c0, s) & chunk_mask(local, 0));
  if (local >= 32) {
    u32x8 c1 = ldg256(base + 32);
    r += __popc(amino_match(c1, s) & chunk_mask(local, 1));
  }
  return r;
}
// Occ at two positions of the SAME block with one pass over its row slices
__device__ __forceinline__ void amino_occ2_same_block(const IndexView& ix, uint32_t blk,
                                                      uint32_t la, uint32_t lb, uint32_t s,
                                                      uint32_t& ra, uint32_t& rb) {
  const char* base = reinterpret_cast<const char*>(ix.blocks + size_t(blk) * AMINO_BLOCK_UINT4);
  uint32_t ms = amino_milestone(ix, blk, s);
  u32x8 c0 = ldg256(base);
  uint32_t m = amino_match(c0, s);
  uint32_t a = ms + __popc(m & chunk_mask(la, 0)), b = ms + __popc(m & chunk_mask(lb, 0));
  if ((la > lb ? la : lb) >= 32) {
    u32x8 c1 = ldg256(base + 32);
    m = amino_match(c1, s);
    a += __popc(m & chunk_mask(la, 1));
    b += __popc(m & chunk_mask(lb, 1));
  }
  ra = a;
  rb = b;
}
__device__ __forceinline__ uint32_t amino_symbol_at(const IndexView& ix, uint32_t pos) {
  uint32_t blk = pos >> 6, local = pos & 63, j = local >> 5, t = local & 31;
  AWRY_CHK(uint64_t(blk) * AMINO_BLOCK_UINT4 + 7 < ix.n_blocks_u4);
  const uint32_t* w =
      reinterpret_cast<const uint32_t*>(ix.blocks + size_t(blk) * AMINO_BLOCK_UINT4) + 8 * j;
  uint32_t s = 0;
#pragma unroll
  for (int p = 0; p < 5; p++) s |= ((__ldg(w + p) >> t) & 1u) << p;
  return s;
}

// ---- alphabet-generic scalar steps (one thread) ------------------------------------------
// Used by the seed-table builder, the single-step API, the amino kernels and the slow paths.

// update_range_with_symbol (fm_index.rs:559-582) for device symbol c; requires sp >= 1.
template <int ALPHA>
__device__ __forceinline__ void lf_update(const IndexView& ix, uint32_t& sp, uint32_t& ep, uint32_t c) {
  uint32_t pa = sp - 1, pb = ep;
  uint32_t ra, rb;
  if (ALPHA == 0) {
    uint32_t ba = pa >> 7, bb = pb >> 7;
    DnaBlockRegs b = dna_load_block(ix, ba);
    ra = dna_occ_in_block(ix, b, ba, pa & 127, c);
    if (bb != ba) b = dna_load_block(ix, bb);
    rb = dna_occ_in_block(ix, b, bb, pb & 127, c);
  } else {
    uint32_t ba = pa >> 6, bb = pb >> 6;
    if (ba == bb) {
      amino_occ2_same_block(ix, ba, pa & 63, pb & 63, c, ra, rb);
    } else {
      ra = amino_occ(ix, pa, c);
      rb = amino_occ(ix, pb, c);
    }
  }
  sp = ix.c_lo[c] + ra;
  ep = ix.c_lo[c] + rb - 1;
}

// backstep (fm_index.rs:585-593): LF(row), or 0 when the row holds '$'
template <int ALPHA>
__device__ __forceinline__ uint32_t lf_backstep(const IndexView& ix, uint32_t row) {
  if (ALPHA == 0) {
    uint32_t blk = row >> 7, local = row & 127;
    DnaBlockRegs b = dna_load_block(ix, blk);
    uint32_t c = dna_code_at(b, local);
    if (c >= DNA_SENTINEL) return 0;
    return ix.c_lo[c] + dna_occ_in_block(ix, b, blk, local, c) - 1;
  } else {
    uint32_t s = amino_symbol_at(ix, row);
    if (s == AMINO_SENTINEL) return 0;
    return ix.c_lo[s] + amino_occ(ix, row, s) - 1;
  }
}

__device__ __forceinline__ bool row_is_sampled(const IndexView& ix, uint32_t row) {
  return ix.sa_pow2 ? (row & (ix.sa_ratio - 1)) == 0 : (row % ix.sa_ratio) == 0;
}
// CompressedSuffixArray::reconstruct_value (compressed_suffix_array.rs:76-106)
__device__ __forceinline__ uint64_t sa_sample(const IndexView& ix, uint32_t row) {
  uint64_t e = ix.sa_pow2 ? (row >> ix.sa_ratio_shift) : (row / ix.sa_ratio);
  uint64_t bit = e * ix.sa_bits;
  uint64_t w = bit >> 6;
  uint32_t s = uint32_t(bit & 63);
  AWRY_CHK(w + (s + ix.sa_bits > 64 ? 1 : 0) < ix.n_sa_words);
  uint64_t v = __ldg(ix.sa_words + w) >> s;
  if (s + ix.sa_bits > 64) v |= __ldg(ix.sa_words + w + 1) << (64 - s);
  return ix.sa_bits >= 64 ? v : (v & ((1ull << ix.sa_bits) - 1));
}

}  // namespace awry
