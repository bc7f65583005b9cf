/*
 * awry_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the batched FM-index search path of the reference
 * crate `awry` 0.3.1 (/root/reference/src).  It exists to CHECK the CUDA path and
 * to time the reference's CPU algorithm on the host cores.  Nothing under
 * awry_b200/ (the product) links, imports or executes it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity pinning (see DESIGN.md "Oracle"): the reference cannot be compiled here
 * (no cargo/rustc), so this restatement is pinned by
 *   (1) the literal vectors the reference's own tests hold (bits_per_element KAT,
 *       compressed_suffix_array.rs:184-201; the symbol code tables, alphabet.rs:169-330),
 *   (2) the properties its tests assert (rank == brute force on random blocks,
 *       bwt.rs:391-505; count/locate == brute-force substring search, fm_index.rs:612-664),
 *   (3) the worked example of SURVEY.md Appendix A (557-byte .awry file, sha256 pinned).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef AWRY_ORACLE_H
#define AWRY_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AWO_NUCLEOTIDE = 0, AWO_AMINO = 1 };

typedef struct awo_index {
  uint64_t version;       /* fm_index.rs:19,53 */
  uint64_t sa_ratio;      /* compressed_suffix_array.rs:15 */
  uint64_t bwt_len;       /* fm_index.rs:51 */
  int alphabet;           /* 0 nucleotide, 1 amino (fm_index_file.rs:165-181) */
  int card;               /* alphabet.rs:87-93: 6 / 22 */
  int n_planes;           /* bwt.rs:29,139: 3 / 5 */
  int n_milestones;       /* bwt.rs:28,138: 8 / 24 */
  size_t block_words;     /* u64 words per block: 3*4+8=20 (160 B) / 5*4+24=44 (352 B) */
  uint64_t n_blocks;      /* bwt.rs:302-304 */
  uint64_t *blocks;       /* reference layout: planes then milestones (fm_index_file.rs:58-68) */
  uint64_t prefix_sums[23];
  uint64_t *sa_words;     /* compressed_suffix_array.rs:13 */
  uint64_t n_sa_words;
  unsigned bits;          /* compressed_suffix_array.rs:124-130 */
  unsigned kmer_len;      /* kmer_lookup_table.rs:19 */
  uint64_t n_seqs;
  uint64_t *seq_starts;   /* sequence_index.rs:10-13 */
  char **headers;
  int owns_arrays;        /* 1 if blocks/sa_words were malloc'd by awo_load */
} awo_index;

typedef struct awo_hit {
  uint64_t seq_idx;   /* sequence_index.rs:33-36 */
  uint64_t local_pos;
} awo_hit;

/* exact work counters, for the algorithmic-bytes figure of SURVEY.md 8(d) */
typedef struct awo_stats {
  uint64_t lf_steps;       /* update_range_with_symbol calls                                 */
  uint64_t block_touches;  /* distinct rank blocks touched by those calls (1 or 2 per call)  */
  uint64_t seeded_steps;   /* LF steps a k-mer-seeded search still runs (L-k per query, with early exit) */
  uint64_t seeded_touches; /* distinct blocks touched by those seeded steps                  */
  uint64_t walk_steps;     /* backstep calls in locate                                       */
  uint64_t hits;
} awo_stats;

/* ---- alphabet.rs ---- */
uint8_t awo_ascii_to_index(int alphabet, uint8_t ascii);
uint8_t awo_index_to_code(int alphabet, uint8_t index);
uint8_t awo_code_to_index(int alphabet, uint8_t code);
char awo_index_to_ascii(int alphabet, uint8_t index);

/* ---- compressed_suffix_array.rs ---- */
unsigned awo_bits_per_element(uint64_t bwt_len);
uint64_t awo_compressed_word_len(uint64_t bwt_len, uint64_t ratio);
void awo_sa_set_value(uint64_t *words, unsigned bits, uint64_t value, uint64_t position);
uint64_t awo_sa_reconstruct(const awo_index *ix, uint64_t position);

/* ---- simd_instructions.rs / bwt.rs ---- */
uint32_t awo_masked_popcount(const uint64_t vec[4], uint64_t local_query_position);
uint64_t awo_block_occurrence(const awo_index *ix, const uint64_t *block, uint64_t local, uint8_t sym);
uint64_t awo_global_occurrence(const awo_index *ix, uint64_t pos, uint8_t sym);
uint8_t awo_symbol_at(const awo_index *ix, uint64_t pos);

/* ---- search.rs / fm_index.rs ---- */
void awo_initial_range(const awo_index *ix, uint8_t sym, uint64_t *sp, uint64_t *ep);
void awo_update_range(const awo_index *ix, uint64_t sp, uint64_t ep, uint8_t sym, uint64_t *nsp,
                      uint64_t *nep);
uint64_t awo_backstep(const awo_index *ix, uint64_t pos);
/* returns 0, or -1 for the inputs on which the reference panics / is UB (empty query, '$'/'#') */
int awo_search_range(const awo_index *ix, const uint8_t *q, uint64_t len, uint64_t *sp, uint64_t *ep,
                     awo_stats *st);
int awo_count_string(const awo_index *ix, const uint8_t *q, uint64_t len, uint64_t *count);
/* hits in BWT-row order (fm_index.rs:521); *hits is malloc'd (caller frees) */
int awo_locate_string(const awo_index *ix, const uint8_t *q, uint64_t len, awo_hit **hits,
                      uint64_t *n_hits, awo_stats *st);
int awo_seq_location(const awo_index *ix, uint64_t loc, awo_hit *out);

/* ---- parallel_count / parallel_locate (fm_index.rs:455-487), pthread stand-in for rayon ---- */
int awo_count_batch(const awo_index *ix, const uint8_t *qbytes, const uint64_t *qoff, uint64_t nq,
                    uint64_t *counts, int n_threads, awo_stats *st);
/* hit_off has nq+1 entries (CSR); *hits malloc'd; per-query slices in BWT-row order, or sorted
 * ascending by (seq_idx, local_pos) when `sorted` is non-zero */
int awo_locate_batch(const awo_index *ix, const uint8_t *qbytes, const uint64_t *qoff, uint64_t nq,
                     uint64_t *hit_off, awo_hit **hits, uint64_t *n_hits, int sorted, int n_threads,
                     awo_stats *st);

/* ---- fm_index_file.rs (read side) ---- */
int awo_load(const char *path, awo_index **out);
/* wrap caller-owned reference-layout arrays (no copy) */
int awo_from_parts(int alphabet, uint64_t sa_ratio, uint64_t bwt_len, unsigned kmer_len,
                   const uint64_t *blocks, const uint64_t *prefix_sums, const uint64_t *sa_words,
                   const uint64_t *seq_starts, uint64_t n_seqs, awo_index **out);
void awo_free(awo_index *ix);
void awo_free_ptr(void *p);
const char *awo_last_error(void);
int awo_hw_threads(void);

#ifdef __cplusplus
}
#endif
#endif
