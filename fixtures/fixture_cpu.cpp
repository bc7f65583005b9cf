// fixture_cpu.cpp -- TEST/BENCH FIXTURE BUILDER (host only). Not part of the product path.
//
// The reference builds its index with FmIndex::new (/root/reference/src/fm_index.rs:142-268)
// on top of libsufr 0.6.2, neither of which can run here (no Rust toolchain).  This file
// restates the reference's single construction pass over a suffix array (fm_index.rs:202-240),
// its k-mer table population (kmer_lookup_table.rs:121-167, incl. its quirk of only visiting
// symbol indices 1..card-3) and its `.awry` v1 writer (fm_index_file.rs:42-106,
// sequence_index.rs:144-152), so that tests and bench.py can create index files the reference
// itself could load.  The suffix array comes from a plain prefix-doubling sorter.
//
// Text model (what libsufr hands the reference, fm_index.rs:148-153,:220-229): records
// upper-cased, joined by the delimiter ('N' / 'X'), followed by one '$'.  bwt_len = that length.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512];
int set_err(const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return -1;
}

// counter-based generator: value at position i depends only on (seed, i), so the CPU and
// GPU fixture builders produce identical texts and queries.
inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t rnd_at(uint64_t seed, uint64_t i) { return mix64(mix64(seed) ^ (i * 0xD1342543DE82EF95ull)); }

const char DNA4[4] = {'A', 'C', 'G', 'T'};
const char AMINO20[21] = "ACDEFGHIKLMNPQRSTVWY";

// alphabet.rs:169-248 (ASCII -> index); same table as the oracle, kept separate on purpose
// (fixtures must not depend on oracle/).
uint8_t ascii_to_index(int alphabet, uint8_t ch) {
  if (ch >= 'a' && ch <= 'z') ch = uint8_t(ch - 'a' + 'A');
  if (ch == '$' || ch == '#') return 0;
  if (alphabet == 0) {
    switch (ch) {
      case 'A': return 1;
      case 'C': return 2;
      case 'G': return 3;
      case 'T':
      case 'U': return 5;
      default: return 4;
    }
  }
  static const char L[23] = "$ACDEFGHIKLMNPQRSTVWXY";
  for (int i = 1; i < 22; i++)
    if (i != 20 && L[i] == char(ch)) return uint8_t(i);
  return 20;
}
// alphabet.rs:250-330 (index -> bit-plane code)
uint8_t index_to_code(int alphabet, uint8_t idx) {
  static const uint8_t D[6] = {0x4, 0x6, 0x5, 0x3, 0x2, 0x1};
  static const uint8_t A[22] = {0x00, 0x0c, 0x17, 0x03, 0x06, 0x1e, 0x1a, 0x1b, 0x19, 0x15, 0x1c,
                                0x1d, 0x08, 0x09, 0x04, 0x13, 0x0a, 0x05, 0x16, 0x01, 0x1f, 0x02};
  return alphabet == 0 ? D[idx] : A[idx];
}

struct Shape {
  int card, planes, milestones;
  size_t block_words;
};
Shape shape_of(int alphabet) {
  return alphabet == 0 ? Shape{6, 3, 8, 20} : Shape{22, 5, 24, 44};
}

unsigned bits_per_element(uint64_t bwt_len) {  // compressed_suffix_array.rs:124-130
  uint64_t v = bwt_len - 1;
  return v ? 64u - unsigned(__builtin_clzll(v)) : 0u;
}
uint64_t sa_word_len(uint64_t bwt_len, uint64_t ratio) {  // compressed_suffix_array.rs:113-123
  unsigned __int128 t = (unsigned __int128)((bwt_len + ratio - 1) / ratio) * bits_per_element(bwt_len);
  return uint64_t((t + 63) / 64);
}

// ---- suffix sorting: prefix doubling with in-place group refinement ----
// sym[i] < 32, sym[n-1] == 0 is the unique smallest symbol.
void suffix_sort(const uint8_t* sym, uint64_t n, std::vector<uint32_t>& sa) {
  const unsigned P = 12;  // 12 symbols x 5 bits in the first-pass key
  struct KP {
    uint64_t key;
    uint32_t pos;
  };
  std::vector<KP> kp(n);
  for (uint64_t s = 0; s < n; s++) {
    uint64_t key = 0;
    for (unsigned j = 0; j < P; j++) key = (key << 5) | (s + j < n ? sym[s + j] : 0);
    kp[s] = KP{key, uint32_t(s)};
  }
  std::sort(kp.begin(), kp.end(), [](const KP& a, const KP& b) { return a.key < b.key; });
  sa.resize(n);
  std::vector<uint32_t> rank(n);
  bool any = false;
  {
    uint64_t gs = 0;
    for (uint64_t i = 0; i < n; i++) {
      if (i > 0 && kp[i].key != kp[i - 1].key) gs = i;
      if (i > 0 && kp[i].key == kp[i - 1].key) any = true;
      sa[i] = kp[i].pos;
      rank[kp[i].pos] = uint32_t(gs);
    }
  }
  std::vector<KP>().swap(kp);
  std::vector<uint32_t> keys;
  for (uint64_t h = P; any; h *= 2) {
    any = false;
    uint64_t i = 0;
    while (i < n) {
      uint32_t r = rank[sa[i]];
      uint64_t j = i + 1;
      while (j < n && rank[sa[j]] == r) j++;
      if (j - i > 1) {
        std::sort(sa.begin() + i, sa.begin() + j,
                  [&](uint32_t a, uint32_t b) { return rank[a + h] < rank[b + h]; });
        keys.resize(j - i);
        for (uint64_t x = i; x < j; x++) keys[x - i] = rank[sa[x] + h];
        uint64_t gs = i;
        for (uint64_t x = i; x < j; x++) {
          if (x > i && keys[x - i] != keys[x - i - 1]) gs = x;
          if (x > i && keys[x - i] == keys[x - i - 1]) any = true;
          rank[sa[x]] = uint32_t(gs);
        }
      }
      i = j;
    }
  }
}

// ---- rank in the reference block layout (for the k-mer table only) ----
struct RefIndex {
  int alphabet;
  Shape sh;
  uint64_t bwt_len;
  const uint64_t* blocks;
  const uint64_t* prefix_sums;
};
uint64_t occ(const RefIndex& ix, uint64_t pos, uint8_t symidx) {  // bwt.rs:338-357 (exact code match)
  const uint64_t* b = ix.blocks + (pos / 256) * ix.sh.block_words;
  uint64_t local = pos % 256;
  uint8_t code = index_to_code(ix.alphabet, symidx);
  uint64_t cnt = b[4 * ix.sh.planes + symidx];
  for (unsigned w = 0; w <= local / 64; w++) {
    uint64_t m = ~0ull;
    for (int p = 0; p < ix.sh.planes; p++) {
      uint64_t v = b[4 * p + w];
      m &= ((code >> p) & 1) ? v : ~v;
    }
    if (w == local / 64) m &= ~0ull >> (63 - local % 64);
    cnt += uint64_t(__builtin_popcountll(m));
  }
  return cnt;
}
void update(const RefIndex& ix, uint64_t& sp, uint64_t& ep, uint8_t s) {  // fm_index.rs:559-582
  uint64_t c = ix.prefix_sums[s];
  uint64_t nsp = c + occ(ix, sp - 1, s);
  uint64_t nep = c + occ(ix, ep, s) - 1;
  sp = nsp;
  ep = nep;
}
// kmer_lookup_table.rs:138-167
void populate_rec(const RefIndex& ix, uint64_t* table, unsigned k, uint64_t sp,
                  uint64_t ep, unsigned cur_len, uint64_t cur_idx, uint64_t mult) {
  if (cur_len == k) {
    table[2 * cur_idx] = sp;
    table[2 * cur_idx + 1] = ep;
    return;
  }
  unsigned enc = unsigned(ix.sh.card - 2);
  for (unsigned idx = 1; idx < enc; idx++) {
    uint64_t nsp = sp, nep = ep;
    update(ix, nsp, nep, uint8_t(idx));
    populate_rec(ix, table, k, nsp, nep, cur_len + 1, cur_idx + idx * mult, mult * enc);
  }
}

uint64_t ipow(uint64_t b, unsigned e) {
  uint64_t r = 1;
  while (e--) r *= b;
  return r;
}

}  // namespace

extern "C" {

const char* fx_last_error(void) { return g_err; }

// alphabet 0: uniform {A,C,G,T}; 1: uniform over the 20 standard amino acids. Writes n bytes.
void fx_gen_text(int alphabet, uint64_t n, uint64_t seed, uint8_t* out) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      for (uint64_t i = lo; i < hi; i++) {
        uint64_t r = rnd_at(seed, i);
        out[i] = alphabet == 0 ? uint8_t(DNA4[r >> 62]) : uint8_t(AMINO20[((r >> 32) * 20) >> 32]);
      }
    });
  for (auto& x : th) x.join();
}

// Exact substrings of text[0..n) at seeded uniform positions; qbytes gets nq*qlen bytes.
// If positions != NULL the start offsets are written there.
void fx_gen_substring_queries(const uint8_t* text, uint64_t n, uint64_t nq, uint64_t qlen,
                              uint64_t seed, uint8_t* qbytes, uint64_t* positions) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  uint64_t span = n - qlen + 1;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      uint64_t lo = nq * t / nt, hi = nq * (t + 1) / nt;
      for (uint64_t q = lo; q < hi; q++) {
        uint64_t r = rnd_at(seed, q);
        uint64_t pos = uint64_t(((unsigned __int128)r * span) >> 64);
        memcpy(qbytes + q * qlen, text + pos, qlen);
        if (positions) positions[q] = pos;
      }
    });
  for (auto& x : th) x.join();
}

// text[pos .. pos+len) of the synthetic text of (alphabet, seed) for each listed position, without
// materialising the text: lets full-size tests verify located positions.
void fx_gen_text_windows(int alphabet, uint64_t seed, const uint64_t* positions, uint64_t n, uint64_t len,
                         uint8_t* out) {
  unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([=] {
      uint64_t lo = n * t / nt, hi = n * (t + 1) / nt;
      for (uint64_t i = lo; i < hi; i++)
        for (uint64_t k = 0; k < len; k++) {
          uint64_t r = rnd_at(seed, positions[i] + k);
          out[i * len + k] = alphabet == 0 ? uint8_t(DNA4[r >> 62]) : uint8_t(AMINO20[((r >> 32) * 20) >> 32]);
        }
    });
  for (auto& x : th) x.join();
}

uint64_t fx_num_blocks(uint64_t bwt_len) { return (bwt_len + 255) / 256; }
uint64_t fx_block_words(int alphabet) { return shape_of(alphabet).block_words; }
uint64_t fx_sa_words(uint64_t bwt_len, uint64_t ratio) { return sa_word_len(bwt_len, ratio); }
uint64_t fx_table_entries(int alphabet, unsigned k) { return ipow(uint64_t(shape_of(alphabet).card - 2), k); }

// Suffix array of text[0..n) + '$' (bwt_len = n+1 rows); sa_out has n+1 u32 entries.
int fx_suffix_array(int alphabet, const uint8_t* text, uint64_t n, uint32_t* sa_out) {
  if (n + 1 >= (1ull << 32)) return set_err("CPU fixture builder is limited to < 2^32 symbols");
  std::vector<uint8_t> sym(n + 1);
  for (uint64_t i = 0; i < n; i++) {
    sym[i] = ascii_to_index(alphabet, text[i]);
    if (sym[i] == 0) return set_err("text contains a sentinel");
  }
  sym[n] = 0;
  std::vector<uint32_t> sa;
  suffix_sort(sym.data(), n + 1, sa);
  memcpy(sa_out, sa.data(), (n + 1) * sizeof(uint32_t));
  return 0;
}

// fm_index.rs:202-240: one pass over the suffix array filling sampled SA, milestones, BWT planes,
// letter counts -> prefix sums.  Outputs (caller-allocated, zero-initialised here):
//   blocks      fx_num_blocks * fx_block_words u64   (planes then milestones per block)
//   prefix_sums card+1 u64
//   sa_words    fx_sa_words u64
int fx_build_parts(int alphabet, const uint8_t* text, uint64_t n, const uint32_t* sa, uint64_t ratio,
                   uint64_t* blocks, uint64_t* prefix_sums, uint64_t* sa_words) {
  if (ratio == 0) return set_err("ratio must be >= 1");
  Shape sh = shape_of(alphabet);
  uint64_t bwt_len = n + 1;
  unsigned bits = bits_per_element(bwt_len);
  memset(blocks, 0, fx_num_blocks(bwt_len) * sh.block_words * 8);
  memset(sa_words, 0, sa_word_len(bwt_len, ratio) * 8);
  std::vector<uint64_t> counts(sh.card, 0);
  for (uint64_t row = 0; row < bwt_len; row++) {
    uint64_t v = sa[row];
    if (row % ratio == 0) {  // compressed_suffix_array.rs:51-64
      unsigned __int128 bp = (unsigned __int128)(row / ratio) * bits;
      uint64_t w = uint64_t(bp / 64);
      unsigned b = unsigned(bp % 64);
      sa_words[w] |= v << b;
      if (b + bits > 64) sa_words[w + 1] |= v >> (64 - b);
    }
    uint64_t* blk = blocks + (row / 256) * sh.block_words;
    if (row % 256 == 0)
      for (int c = 0; c < sh.card; c++) blk[4 * sh.planes + c] = counts[c];
    uint8_t idx = v == 0 ? 0 : ascii_to_index(alphabet, text[v - 1]);
    uint8_t code = index_to_code(alphabet, idx);
    uint64_t local = row % 256;
    for (int p = 0; p < sh.planes; p++)
      if ((code >> p) & 1) blk[4 * p + local / 64] |= 1ull << (local % 64);
    counts[idx]++;
  }
  uint64_t acc = 0;
  for (int c = 0; c <= sh.card; c++) {
    prefix_sums[c] = acc;
    if (c < sh.card) acc += counts[c];
  }
  return 0;
}

// kmer_lookup_table.rs:121-167: fills `table` (2 u64 per entry: start,end) exactly as the
// reference does -- entries it never visits stay SearchRange::zero() = (1,0).
void fx_populate_kmer_table(int alphabet, uint64_t bwt_len, const uint64_t* blocks,
                            const uint64_t* prefix_sums, unsigned k, uint64_t* table) {
  RefIndex ix{alphabet, shape_of(alphabet), bwt_len, blocks, prefix_sums};
  uint64_t n_entries = fx_table_entries(alphabet, k);
  for (uint64_t i = 0; i < n_entries; i++) {
    table[2 * i] = 1;
    table[2 * i + 1] = 0;
  }
  unsigned enc = unsigned(ix.sh.card - 2);
  for (unsigned s = 1; s < enc; s++) {
    uint64_t sp = prefix_sums[s], ep = prefix_sums[s + 1] - 1;  // search.rs:43-48
    populate_rec(ix, table, k, sp, ep, 1, s, enc);
  }
}

// fm_index_file.rs:42-106 + sequence_index.rs:144-152.  table may be NULL: then every entry is
// written as (1,0) (what the search path ignores anyway).  headers: n_seqs C strings.
int fx_write_awry(const char* path, int alphabet, uint64_t ratio, uint64_t bwt_len,
                  const uint64_t* blocks, const uint64_t* prefix_sums, const uint64_t* sa_words,
                  unsigned k, const uint64_t* table, const uint64_t* seq_starts,
                  const char* const* headers, uint64_t n_seqs) {
  Shape sh = shape_of(alphabet);
  FILE* f = fopen(path, "wb");
  if (!f) return set_err("cannot open output file");
  bool ok = true;
  auto put = [&](const void* p, size_t nbytes) {
    if (ok && nbytes && fwrite(p, 1, nbytes, f) != nbytes) ok = false;
  };
  put("AWRY-Index\n", 11);
  uint64_t hdr[4] = {1, ratio, bwt_len, uint64_t(alphabet)};
  put(hdr, sizeof hdr);
  put(blocks, fx_num_blocks(bwt_len) * sh.block_words * 8);
  put(prefix_sums, size_t(sh.card + 1) * 8);
  put(sa_words, sa_word_len(bwt_len, ratio) * 8);
  uint8_t kb = uint8_t(k);
  put(&kb, 1);
  uint64_t n_entries = fx_table_entries(alphabet, k);
  if (table) {
    put(table, n_entries * 16);
  } else {
    std::vector<uint64_t> chunk(2 * 65536);
    for (size_t i = 0; i < chunk.size(); i += 2) {
      chunk[i] = 1;
      chunk[i + 1] = 0;
    }
    for (uint64_t done = 0; done < n_entries;) {
      uint64_t m = std::min<uint64_t>(65536, n_entries - done);
      put(chunk.data(), m * 16);
      done += m;
    }
  }
  put(&n_seqs, 8);
  for (uint64_t i = 0; i < n_seqs; i++) {
    uint64_t hl = strlen(headers[i]);
    put(&seq_starts[i], 8);
    put(&hl, 8);
    put(headers[i], hl);
  }
  if (fclose(f) != 0) ok = false;
  return ok ? 0 : set_err("write failed");
}

}  // extern "C"
