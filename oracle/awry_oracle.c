/*
 * awry_oracle.c -- TEST INFRASTRUCTURE ONLY (see awry_oracle.h).
 *
 * Plain-C restatement of the reference's CPU search path, in the reference's own
 * in-memory block layout (160-B nucleotide / 352-B amino blocks) and `.awry` v1 file
 * format.  Scalar 4 x u64 plane arithmetic stands in for the AVX2/NEON intrinsics of
 * simd_instructions.rs (bit-identical by construction; gcc -O3 -mavx2 vectorises it).
 *
 * Pinning status: the reference is a Rust crate and no Rust toolchain exists in the build image, so
 * no output of the reference binary itself is available (no oracle/_ref).  This restatement is pinned by
 * what the reference's own tests hold for the path -- its literal vectors (bits_per_element table,
 * symbol code tables) and its properties re-run at the same sizes (rank == brute force, sampled-SA round
 * trip, every k-mer counted and located == substring search, save -> load equality) -- and by the worked
 * example of SURVEY.md Appendix A (tests/test_oracle.py, tests/golden/appendix_a.awry).  It is NOT pinned
 * by reference-generated vectors: by the task's rule, "parity unpinned" with respect to the reference
 * binary.
 */
#define _GNU_SOURCE
#include "awry_oracle.h"

#include <errno.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static __thread char g_err[512];
static int set_err(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return -1;
}
const char *awo_last_error(void) { return g_err; }
void awo_free_ptr(void *p) { free(p); }
int awo_hw_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------ alphabet.rs */

static const char AMINO_LETTERS[22] = {'$', 'A', 'C', 'D', 'E', 'F', 'G', 'H', 'I', 'K', 'L',
                                       'M', 'N', 'P', 'Q', 'R', 'S', 'T', 'V', 'W', 'X', 'Y'};
/* alphabet.rs:255-303 (index -> strided bit-vector code) */
static const uint8_t AMINO_CODES[22] = {0x00, 0x0c, 0x17, 0x03, 0x06, 0x1e, 0x1a, 0x1b,
                                        0x19, 0x15, 0x1c, 0x1d, 0x08, 0x09, 0x04, 0x13,
                                        0x0a, 0x05, 0x16, 0x01, 0x1f, 0x02};
/* alphabet.rs:319-327 */
static const uint8_t DNA_CODES[6] = {0x4, 0x6, 0x5, 0x3, 0x2, 0x1};
static const char DNA_LETTERS[6] = {'$', 'A', 'C', 'G', 'N', 'T'};

/* alphabet.rs:109-114 (to_ascii_uppercase) + :169-248 (to_index, Ascii arm) */
uint8_t awo_ascii_to_index(int alphabet, uint8_t ch) {
  if (ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 'a' + 'A');
  if (ch == '#' || ch == '$') return 0;
  if (alphabet == AWO_NUCLEOTIDE) {
    switch (ch) {
      case 'A': return 1;
      case 'C': return 2;
      case 'G': return 3;
      case 'T':
      case 'U': return 5;
      default: return 4; /* ambiguity char N */
    }
  }
  for (int i = 1; i < 22; i++)
    if (i != 20 && AMINO_LETTERS[i] == (char)ch) return (uint8_t)i;
  return 20; /* ambiguity character X */
}

/* alphabet.rs:250-330 (to_bit_vector, Index arm) */
uint8_t awo_index_to_code(int alphabet, uint8_t index) {
  if (alphabet == AWO_NUCLEOTIDE) return index < 6 ? DNA_CODES[index] : 0x2;
  return index < 22 ? AMINO_CODES[index] : 0x1f;
}

/* alphabet.rs:197-222, :237-244 (to_index, BitVector arm): unknown codes decode to X / N */
uint8_t awo_code_to_index(int alphabet, uint8_t code) {
  if (alphabet == AWO_NUCLEOTIDE) {
    switch (code) {
      case 0x4: return 0;
      case 0x6: return 1;
      case 0x5: return 2;
      case 0x3: return 3;
      case 0x1: return 5;
      default: return 4;
    }
  }
  for (int i = 0; i < 22; i++)
    if (i != 20 && AMINO_CODES[i] == code) return (uint8_t)i;
  return 20;
}

/* alphabet.rs:333-409 (to_ascii, Index arm) */
char awo_index_to_ascii(int alphabet, uint8_t index) {
  if (alphabet == AWO_NUCLEOTIDE) return index < 6 ? DNA_LETTERS[index] : 'N';
  return index < 22 ? AMINO_LETTERS[index] : 'X';
}

/* ------------------------------------------------------------------ compressed_suffix_array.rs */

/* compressed_suffix_array.rs:124-130 */
unsigned awo_bits_per_element(uint64_t bwt_len) {
  uint64_t largest = bwt_len - 1;
  unsigned lz = largest ? (unsigned)__builtin_clzll(largest) : 64u;
  return 64u - lz;
}

/* compressed_suffix_array.rs:113-123 */
uint64_t awo_compressed_word_len(uint64_t bwt_len, uint64_t ratio) {
  unsigned bits = awo_bits_per_element(bwt_len);
  uint64_t n_elems = (bwt_len + ratio - 1) / ratio;
  unsigned __int128 total = (unsigned __int128)n_elems * bits;
  return (uint64_t)((total + 63) / 64);
}

/* compressed_suffix_array.rs:51-64 */
void awo_sa_set_value(uint64_t *words, unsigned bits, uint64_t value, uint64_t position) {
  unsigned __int128 bitpos = (unsigned __int128)position * bits;
  uint64_t w = (uint64_t)(bitpos / 64);
  unsigned b = (unsigned)(bitpos % 64);
  words[w] |= value << b;
  if (b + bits > 64) words[w + 1] |= (b == 0) ? 0 : (value >> (64 - b));
}

/* compressed_suffix_array.rs:76-106 (position is the UNsampled BWT row, must be sampled) */
uint64_t awo_sa_reconstruct(const awo_index *ix, uint64_t position) {
  uint64_t sampled = position / ix->sa_ratio;
  unsigned __int128 bitpos = (unsigned __int128)sampled * ix->bits;
  uint64_t w = (uint64_t)(bitpos / 64);
  unsigned first_start = (unsigned)(bitpos % 64);
  unsigned first_n = ix->bits < 64 - first_start ? ix->bits : 64 - first_start;
  unsigned second_n = ix->bits - first_n;
  uint64_t first_mask = first_n >= 64 ? ~0ull : ((1ull << first_n) - 1);
  uint64_t v = (ix->sa_words[w] >> first_start) & first_mask;
  if (second_n != 0) {
    uint64_t second_mask = (1ull << second_n) - 1;
    v |= (ix->sa_words[w + 1] & second_mask) << first_n;
  }
  return v;
}

/* ------------------------------------------------------------------ simd_instructions.rs */

typedef struct {
  uint64_t w[4];
} v256; /* simd_instructions.rs:35-37 */

static inline v256 v_load(const uint64_t *p) { /* :44-48 */
  v256 r;
  for (int i = 0; i < 4; i++) r.w[i] = p[i];
  return r;
}
static inline v256 v_and(v256 a, v256 b) { /* :78-82 */
  for (int i = 0; i < 4; i++) a.w[i] &= b.w[i];
  return a;
}
static inline v256 v_or(v256 a, v256 b) { /* :84-88 */
  for (int i = 0; i < 4; i++) a.w[i] |= b.w[i];
  return a;
}
static inline v256 v_andnot(v256 a, v256 b) { /* :90-94  (~a) & b */
  for (int i = 0; i < 4; i++) a.w[i] = ~a.w[i] & b.w[i];
  return a;
}

/* simd_instructions.rs:96-121 -- popcount of bits 0..=pos (inclusive) */
static inline uint32_t v_masked_popcount(v256 v, uint64_t pos) {
  uint64_t masks[4] = {0, 0, 0, 0};
  unsigned qw = (unsigned)(pos / 64);
  for (unsigned i = 0; i < qw; i++) masks[i] = ~0ull;
  masks[qw] = ~0ull >> (63 - (pos % 64));
  uint32_t pc = 0;
  for (int i = 0; i < 4; i++) pc += (uint32_t)__builtin_popcountll(v.w[i] & masks[i]);
  return pc;
}
uint32_t awo_masked_popcount(const uint64_t vec[4], uint64_t pos) {
  return v_masked_popcount(v_load(vec), pos);
}

/* ------------------------------------------------------------------ bwt.rs */

/* bwt.rs:114-135 */
static inline uint64_t dna_block_occ(const uint64_t *blk, uint64_t local, uint8_t sym) {
  uint64_t milestone = blk[12 + sym]; /* :96-101, planes are 3*4 words */
  v256 v0 = v_load(blk), v1 = v_load(blk + 4), v2 = v_load(blk + 8), occ;
  switch (sym) {
    case 1: occ = v_and(v1, v2); break;                  /* A 0b110 */
    case 2: occ = v_and(v0, v2); break;                  /* C 0b101 */
    case 3: occ = v_and(v0, v1); break;                  /* G 0b011 */
    case 4: occ = v_andnot(v2, v_andnot(v0, v1)); break; /* N 0b010 */
    case 5: occ = v_andnot(v2, v_andnot(v1, v0)); break; /* T 0b001 */
    default: abort();                                    /* reference panics (:127) */
  }
  return milestone + v_masked_popcount(occ, local);
}

/* bwt.rs:230-271 */
static inline uint64_t amino_block_occ(const uint64_t *blk, uint64_t local, uint8_t sym) {
  uint64_t milestone = blk[20 + sym];
  v256 v0 = v_load(blk), v1 = v_load(blk + 4), v2 = v_load(blk + 8), v3 = v_load(blk + 12),
       v4 = v_load(blk + 16), o;
  switch (sym) {
    case 1: o = v_and(v2, v_andnot(v4, v3)); break;
    case 2: o = v_andnot(v3, v_and(v_and(v0, v1), v2)); break;
    case 3: o = v_andnot(v4, v_and(v0, v1)); break;
    case 4: o = v_andnot(v4, v_and(v1, v2)); break;
    case 5: o = v_andnot(v0, v_and(v_and(v1, v2), v3)); break;
    case 6: o = v_andnot(v2, v_andnot(v0, v4)); break;
    case 7: o = v_andnot(v2, v_and(v0, v_and(v1, v3))); break;
    case 8: o = v_andnot(v2, v_andnot(v1, v4)); break;
    case 9: o = v_andnot(v1, v_andnot(v3, v4)); break;
    case 10: o = v_andnot(v1, v_andnot(v0, v4)); break;
    case 11: o = v_andnot(v1, v_and(v3, v_and(v2, v0))); break;
    case 12: o = v_andnot(v_or(v0, v1), v_andnot(v2, v3)); break;
    case 13: o = v_and(v3, v_andnot(v4, v0)); break;
    case 14: o = v_andnot(v_or(v0, v1), v_andnot(v3, v2)); break;
    case 15: o = v_andnot(v2, v_andnot(v3, v4)); break;
    case 16: o = v_and(v1, v_andnot(v4, v3)); break;
    case 17: o = v_and(v0, v_andnot(v4, v2)); break;
    case 18: o = v_andnot(v3, v_andnot(v0, v4)); break;
    case 19: o = v_andnot(v_or(v1, v2), v_andnot(v3, v0)); break;
    case 20: o = v_and(v_and(v0, v1), v_and(v2, v3)); break;
    case 21: o = v_andnot(v_or(v0, v2), v_andnot(v3, v1)); break;
    default: abort(); /* reference panics (:263) */
  }
  return milestone + v_masked_popcount(o, local);
}

uint64_t awo_block_occurrence(const awo_index *ix, const uint64_t *block, uint64_t local,
                              uint8_t sym) {
  return ix->alphabet == AWO_NUCLEOTIDE ? dna_block_occ(block, local, sym)
                                        : amino_block_occ(block, local, sym);
}

/* bwt.rs:338-357 */
uint64_t awo_global_occurrence(const awo_index *ix, uint64_t pos, uint8_t sym) {
  uint64_t block_idx = pos / 256, local = pos % 256;
  const uint64_t *blk = ix->blocks + block_idx * ix->block_words;
  return ix->alphabet == AWO_NUCLEOTIDE ? dna_block_occ(blk, local, sym)
                                        : amino_block_occ(blk, local, sym);
}

/* bwt.rs:307-325 + block symbol_at :53-62 / :162-174 + extract_bit simd_instructions.rs:57-63 */
uint8_t awo_symbol_at(const awo_index *ix, uint64_t pos) {
  uint64_t block_idx = pos / 256, local = pos % 256;
  const uint64_t *blk = ix->blocks + block_idx * ix->block_words;
  uint8_t code = 0;
  for (int bit = 0; bit < ix->n_planes; bit++)
    code |= (uint8_t)(((blk[4 * bit + local / 64] >> (local % 64)) & 1) << bit);
  return awo_code_to_index(ix->alphabet, code);
}

/* ------------------------------------------------------------------ search.rs / fm_index.rs */

/* search.rs:43-48 */
void awo_initial_range(const awo_index *ix, uint8_t sym, uint64_t *sp, uint64_t *ep) {
  *sp = ix->prefix_sums[sym];
  *ep = ix->prefix_sums[sym + 1] - 1;
}

/* fm_index.rs:559-582 */
void awo_update_range(const awo_index *ix, uint64_t sp, uint64_t ep, uint8_t sym, uint64_t *nsp,
                      uint64_t *nep) {
  uint64_t c = ix->prefix_sums[sym];
  *nsp = c + awo_global_occurrence(ix, sp - 1, sym);
  *nep = c + awo_global_occurrence(ix, ep, sym) - 1;
}

/* fm_index.rs:585-593 */
uint64_t awo_backstep(const awo_index *ix, uint64_t pos) {
  uint8_t sym = awo_symbol_at(ix, pos);
  if (sym == 0) return 0;
  return ix->prefix_sums[sym] + awo_global_occurrence(ix, pos, sym) - 1;
}

static inline void count_step(awo_stats *st, uint64_t sp, uint64_t ep, int seeded) {
  if (!st) return;
  uint64_t d = ((sp - 1) / 256 == ep / 256) ? 1 : 2;
  st->lf_steps++;
  st->block_touches += d;
  if (seeded) {
    st->seeded_steps++;
    st->seeded_touches += d;
  }
}

/* fm_index.rs:402-438, with kmer_lookup_table.rs:90-110 inlined for the len >= k arm.
 * Returns -1 where the reference panics or is undefined. */
int awo_search_range(const awo_index *ix, const uint8_t *q, uint64_t len, uint64_t *sp_out,
                     uint64_t *ep_out, awo_stats *st) {
  if (len == 0) return set_err("empty query (reference: unwrap on None, fm_index.rs:406)");
  for (uint64_t i = 0; i < len; i++)
    if (q[i] == '$' || q[i] == '#')
      return set_err("query contains a sentinel (reference: panic, bwt.rs:127)");
  uint64_t sp, ep;
  uint64_t k = ix->kmer_len;
  if (len < k) {
    /* fm_index.rs:403-418 */
    awo_initial_range(ix, awo_ascii_to_index(ix->alphabet, q[len - 1]), &sp, &ep);
    for (uint64_t i = len - 1; i-- > 0;) {
      if (sp > ep) break;
      count_step(st, sp, ep, 1);
      awo_update_range(ix, sp, ep, awo_ascii_to_index(ix->alphabet, q[i]), &sp, &ep);
    }
  } else {
    /* kmer_lookup_table.rs:90-110: k-1 updates, no table read, no emptiness check */
    awo_initial_range(ix, awo_ascii_to_index(ix->alphabet, q[len - 1]), &sp, &ep);
    uint64_t i = len - 1;
    for (uint64_t t = 0; t + 1 < k && i > 0; t++) {
      i--;
      count_step(st, sp, ep, 0);
      awo_update_range(ix, sp, ep, awo_ascii_to_index(ix->alphabet, q[i]), &sp, &ep);
    }
    /* fm_index.rs:421-434: the rest, with early break */
    while (i > 0) {
      i--;
      if (sp > ep) break;
      count_step(st, sp, ep, 1);
      awo_update_range(ix, sp, ep, awo_ascii_to_index(ix->alphabet, q[i]), &sp, &ep);
    }
  }
  *sp_out = sp;
  *ep_out = ep;
  return 0;
}

/* fm_index.rs:499-501 + search.rs:66-71 */
int awo_count_string(const awo_index *ix, const uint8_t *q, uint64_t len, uint64_t *count) {
  uint64_t sp, ep;
  if (awo_search_range(ix, q, len, &sp, &ep, NULL)) return -1;
  *count = sp > ep ? 0 : ep - sp + 1;
  return 0;
}

/* sequence_index.rs:108-141, INTENDED semantics: the record with the largest start <= loc.
 * (The reference recursion does not terminate for most positions of multi-record inputs --
 * SURVEY.md Q4 -- and returns (0, loc - start[0]) for single-record inputs, which this matches.) */
int awo_seq_location(const awo_index *ix, uint64_t loc, awo_hit *out) {
  if (ix->n_seqs == 0) return set_err("empty sequence index");
  uint64_t lo = 0, hi = ix->n_seqs - 1;
  while (lo < hi) {
    uint64_t mid = (lo + hi + 1) / 2;
    if (ix->seq_starts[mid] <= loc)
      lo = mid;
    else
      hi = mid - 1;
  }
  out->seq_idx = lo;
  out->local_pos = loc - ix->seq_starts[lo];
  return 0;
}

/* fm_index.rs:516-544 */
int awo_locate_string(const awo_index *ix, const uint8_t *q, uint64_t len, awo_hit **hits_out,
                      uint64_t *n_hits, awo_stats *st) {
  uint64_t sp, ep;
  *hits_out = NULL;
  *n_hits = 0;
  if (awo_search_range(ix, q, len, &sp, &ep, st)) return -1;
  if (sp > ep) return 0;
  uint64_t n = ep - sp + 1;
  awo_hit *hits = (awo_hit *)malloc(n * sizeof(awo_hit));
  if (!hits) return set_err("out of memory for %llu hits", (unsigned long long)n);
  for (uint64_t row = sp; row <= ep; row++) {
    uint64_t steps = 0, j = row;
    while (j % ix->sa_ratio != 0) { /* compressed_suffix_array.rs:109-111 */
      j = awo_backstep(ix, j);
      steps++;
    }
    uint64_t loc = (awo_sa_reconstruct(ix, j) + steps) % ix->bwt_len;
    awo_seq_location(ix, loc, &hits[row - sp]);
    if (st) st->walk_steps += steps;
  }
  if (st) st->hits += n;
  *hits_out = hits;
  *n_hits = n;
  return 0;
}

/* ------------------------------------------------------------------ batched (rayon stand-in) */

typedef struct {
  const awo_index *ix;
  const uint8_t *qbytes;
  const uint64_t *qoff;
  uint64_t nq;
  uint64_t *counts;    /* count mode */
  awo_hit **per_query; /* locate mode */
  uint64_t *per_query_n;
  int locate, sorted;
  volatile uint64_t next;
  volatile int failed;
  pthread_mutex_t mu;
  awo_stats total;
  char err[256];
} batch_job;

static int hit_cmp(const void *a, const void *b) {
  const awo_hit *x = (const awo_hit *)a, *y = (const awo_hit *)b;
  if (x->seq_idx != y->seq_idx) return x->seq_idx < y->seq_idx ? -1 : 1;
  if (x->local_pos != y->local_pos) return x->local_pos < y->local_pos ? -1 : 1;
  return 0;
}

static void *batch_worker(void *arg) {
  batch_job *job = (batch_job *)arg;
  awo_stats st;
  memset(&st, 0, sizeof st);
  const uint64_t CHUNK = 256; /* dynamic chunking: stand-in for rayon work stealing */
  for (;;) {
    uint64_t lo = __atomic_fetch_add(&job->next, CHUNK, __ATOMIC_RELAXED);
    if (lo >= job->nq || job->failed) break;
    uint64_t hi = lo + CHUNK < job->nq ? lo + CHUNK : job->nq;
    for (uint64_t i = lo; i < hi; i++) {
      const uint8_t *q = job->qbytes + job->qoff[i];
      uint64_t len = job->qoff[i + 1] - job->qoff[i];
      int rc;
      if (!job->locate) {
        uint64_t sp, ep;
        rc = awo_search_range(job->ix, q, len, &sp, &ep, &st);
        if (!rc) job->counts[i] = sp > ep ? 0 : ep - sp + 1;
      } else {
        rc = awo_locate_string(job->ix, q, len, &job->per_query[i], &job->per_query_n[i], &st);
        if (!rc && job->sorted && job->per_query_n[i] > 1)
          qsort(job->per_query[i], job->per_query_n[i], sizeof(awo_hit), hit_cmp);
      }
      if (rc) {
        pthread_mutex_lock(&job->mu);
        if (!job->failed) {
          job->failed = 1;
          snprintf(job->err, sizeof job->err, "query %llu: %s", (unsigned long long)i, g_err);
        }
        pthread_mutex_unlock(&job->mu);
        break;
      }
    }
  }
  pthread_mutex_lock(&job->mu);
  job->total.lf_steps += st.lf_steps;
  job->total.block_touches += st.block_touches;
  job->total.seeded_steps += st.seeded_steps;
  job->total.seeded_touches += st.seeded_touches;
  job->total.walk_steps += st.walk_steps;
  job->total.hits += st.hits;
  pthread_mutex_unlock(&job->mu);
  return NULL;
}

static int run_batch(batch_job *job, int n_threads) {
  if (n_threads <= 0) n_threads = awo_hw_threads();
  if ((uint64_t)n_threads > job->nq) n_threads = job->nq ? (int)job->nq : 1;
  pthread_mutex_init(&job->mu, NULL);
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  int started = 0;
  for (int t = 1; t < n_threads; t++)
    if (pthread_create(&th[started], NULL, batch_worker, job) == 0) started++;
  batch_worker(job);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  free(th);
  pthread_mutex_destroy(&job->mu);
  if (job->failed) return set_err("%s", job->err);
  return 0;
}

/* fm_index.rs:455-460 */
int awo_count_batch(const awo_index *ix, const uint8_t *qbytes, const uint64_t *qoff, uint64_t nq,
                    uint64_t *counts, int n_threads, awo_stats *st) {
  batch_job job;
  memset(&job, 0, sizeof job);
  job.ix = ix;
  job.qbytes = qbytes;
  job.qoff = qoff;
  job.nq = nq;
  job.counts = counts;
  int rc = run_batch(&job, n_threads);
  if (st) *st = job.total;
  return rc;
}

/* fm_index.rs:479-487: per-query vectors, then flattened to CSR for the C interface */
int awo_locate_batch(const awo_index *ix, const uint8_t *qbytes, const uint64_t *qoff, uint64_t nq,
                     uint64_t *hit_off, awo_hit **hits_out, uint64_t *n_hits, int sorted,
                     int n_threads, awo_stats *st) {
  batch_job job;
  memset(&job, 0, sizeof job);
  job.ix = ix;
  job.qbytes = qbytes;
  job.qoff = qoff;
  job.nq = nq;
  job.locate = 1;
  job.sorted = sorted;
  job.per_query = (awo_hit **)calloc(nq ? nq : 1, sizeof(awo_hit *));
  job.per_query_n = (uint64_t *)calloc(nq ? nq : 1, sizeof(uint64_t));
  int rc = run_batch(&job, n_threads);
  uint64_t total = 0;
  if (!rc) {
    for (uint64_t i = 0; i < nq; i++) {
      hit_off[i] = total;
      total += job.per_query_n[i];
    }
    hit_off[nq] = total;
    awo_hit *flat = (awo_hit *)malloc((total ? total : 1) * sizeof(awo_hit));
    for (uint64_t i = 0; i < nq; i++)
      if (job.per_query_n[i])
        memcpy(flat + hit_off[i], job.per_query[i], job.per_query_n[i] * sizeof(awo_hit));
    *hits_out = flat;
    *n_hits = total;
  }
  for (uint64_t i = 0; i < nq; i++) free(job.per_query[i]);
  free(job.per_query);
  free(job.per_query_n);
  if (st) *st = job.total;
  return rc;
}

/* ------------------------------------------------------------------ fm_index_file.rs (read side) */

static void fill_shape(awo_index *ix) {
  if (ix->alphabet == AWO_NUCLEOTIDE) {
    ix->card = 6;
    ix->n_planes = 3;
    ix->n_milestones = 8;
  } else {
    ix->card = 22;
    ix->n_planes = 5;
    ix->n_milestones = 24;
  }
  ix->block_words = (size_t)ix->n_planes * 4 + (size_t)ix->n_milestones;
  ix->n_blocks = (ix->bwt_len + 255) / 256; /* bwt.rs:302-304 */
  ix->bits = awo_bits_per_element(ix->bwt_len);
  ix->n_sa_words = awo_compressed_word_len(ix->bwt_len, ix->sa_ratio);
}

static int read_exact(FILE *f, void *dst, size_t n, const char *what) {
  if (n && fread(dst, 1, n, f) != n) return set_err("short read in %s", what);
  return 0;
}

static uint64_t ipow(uint64_t b, unsigned e) {
  uint64_t r = 1;
  while (e--) r *= b;
  return r;
}

/* fm_index_file.rs:132-160 + :184-287; kmer_lookup_table.rs:55-77; sequence_index.rs:155-183 */
int awo_load(const char *path, awo_index **out) {
  FILE *f = fopen(path, "rb");
  if (!f) return set_err("cannot open %s: %s", path, strerror(errno));
  awo_index *ix = (awo_index *)calloc(1, sizeof *ix);
  ix->owns_arrays = 1;
  char label[11];
  uint64_t hdr[4];
  int rc = read_exact(f, label, 11, "label");
  if (!rc && memcmp(label, "AWRY-Index\n", 11) != 0)
    rc = set_err("file did not start with expected label"); /* :145-150 */
  if (!rc) rc = read_exact(f, hdr, sizeof hdr, "header");
  if (!rc) {
    ix->version = hdr[0]; /* read, not validated (:185-186) */
    ix->sa_ratio = hdr[1];
    ix->bwt_len = hdr[2];
    if (hdr[3] > 1) rc = set_err("invalid symbol alphabet %llu", (unsigned long long)hdr[3]);
    ix->alphabet = (int)hdr[3];
    if (!rc && (ix->sa_ratio == 0 || ix->bwt_len < 2)) rc = set_err("corrupt header");
  }
  if (!rc) {
    fill_shape(ix);
    size_t bytes = ix->n_blocks * ix->block_words * 8;
    if (posix_memalign((void **)&ix->blocks, 64, bytes ? bytes : 64)) rc = set_err("oom (blocks)");
    if (!rc) rc = read_exact(f, ix->blocks, bytes, "bwt blocks");
  }
  if (!rc) rc = read_exact(f, ix->prefix_sums, (size_t)(ix->card + 1) * 8, "prefix sums");
  if (!rc) {
    ix->sa_words = (uint64_t *)malloc((ix->n_sa_words + 1) * 8);
    if (!ix->sa_words) rc = set_err("oom (sa)");
    if (!rc) rc = read_exact(f, ix->sa_words, ix->n_sa_words * 8, "sampled suffix array");
  }
  if (!rc) {
    uint8_t k;
    rc = read_exact(f, &k, 1, "kmer length");
    ix->kmer_len = k;
    /* the table contents are never read by the search path (kmer_lookup_table.rs:90-110): skip */
    uint64_t n_entries = ipow((uint64_t)(ix->card - 2), k);
    if (!rc && fseeko(f, (off_t)(n_entries * 16), SEEK_CUR)) rc = set_err("seek past kmer table");
  }
  if (!rc) {
    uint64_t n;
    rc = read_exact(f, &n, 8, "sequence count");
    if (!rc) {
      ix->n_seqs = n;
      ix->seq_starts = (uint64_t *)calloc(n ? n : 1, 8);
      ix->headers = (char **)calloc(n ? n : 1, sizeof(char *));
      for (uint64_t i = 0; i < n && !rc; i++) {
        uint64_t hl;
        rc = read_exact(f, &ix->seq_starts[i], 8, "sequence start");
        if (!rc) rc = read_exact(f, &hl, 8, "header length");
        if (!rc) {
          ix->headers[i] = (char *)calloc(hl + 1, 1);
          rc = read_exact(f, ix->headers[i], hl, "header");
        }
      }
    }
  }
  fclose(f);
  if (rc) {
    awo_free(ix);
    return rc;
  }
  *out = ix;
  return 0;
}

int awo_from_parts(int alphabet, uint64_t sa_ratio, uint64_t bwt_len, unsigned kmer_len,
                   const uint64_t *blocks, const uint64_t *prefix_sums, const uint64_t *sa_words,
                   const uint64_t *seq_starts, uint64_t n_seqs, awo_index **out) {
  if (alphabet < 0 || alphabet > 1 || sa_ratio == 0 || bwt_len < 2) return set_err("bad parts");
  awo_index *ix = (awo_index *)calloc(1, sizeof *ix);
  ix->version = 1;
  ix->alphabet = alphabet;
  ix->sa_ratio = sa_ratio;
  ix->bwt_len = bwt_len;
  ix->kmer_len = kmer_len;
  fill_shape(ix);
  ix->blocks = (uint64_t *)blocks;
  ix->sa_words = (uint64_t *)sa_words;
  memcpy(ix->prefix_sums, prefix_sums, (size_t)(ix->card + 1) * 8);
  ix->n_seqs = n_seqs ? n_seqs : 1;
  ix->seq_starts = (uint64_t *)calloc(ix->n_seqs, 8);
  if (n_seqs) memcpy(ix->seq_starts, seq_starts, n_seqs * 8);
  ix->owns_arrays = 0;
  *out = ix;
  return 0;
}

void awo_free(awo_index *ix) {
  if (!ix) return;
  if (ix->owns_arrays) {
    free(ix->blocks);
    free(ix->sa_words);
  }
  if (ix->headers)
    for (uint64_t i = 0; i < ix->n_seqs; i++) free(ix->headers[i]);
  free(ix->headers);
  free(ix->seq_starts);
  free(ix);
}
