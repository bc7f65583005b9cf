// awry_b200.hpp -- header-only C++17 host mirror of the reference's Rust API over the C ABI.
//
// The reference is a Rust crate; no Rust toolchain exists in the build image, so this is the
// compiled-language host side that sits above include/awry_b200.h (the Rust facade under rust/
// is the same thing in the reference's own language).  Names, argument meaning and error
// behaviour follow /root/reference/src/fm_index.rs: where the reference returns io::Error this
// throws awry::Error; where it panics (empty query, sentinel in a query) this throws too.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "awry_b200.h"

namespace awry {

struct Error : std::runtime_error {
  int code;
  Error(int c, const char* msg) : std::runtime_error(msg), code(c) {}
};

enum class SymbolAlphabet : uint32_t { Nucleotide = AWRY_NUCLEOTIDE, Amino = AWRY_AMINO };  // alphabet.rs:28-31

struct SearchRange {  // search.rs:25-81
  uint64_t start_ptr = 1, end_ptr = 0;
  static SearchRange zero() { return {1, 0}; }
  bool is_empty() const { return start_ptr > end_ptr; }
  uint64_t len() const { return is_empty() ? 0 : end_ptr - start_ptr + 1; }
};

struct LocalizedSequencePosition {  // sequence_index.rs:32-78
  uint64_t sequence_idx_ = 0, local_position_ = 0;
  uint64_t sequence_idx() const { return sequence_idx_; }
  uint64_t local_position() const { return local_position_; }
  bool operator==(const LocalizedSequencePosition& o) const {
    return sequence_idx_ == o.sequence_idx_ && local_position_ == o.local_position_;
  }
  bool operator<(const LocalizedSequencePosition& o) const {
    return sequence_idx_ != o.sequence_idx_ ? sequence_idx_ < o.sequence_idx_ : local_position_ < o.local_position_;
  }
};

struct FmBuildArgs {  // fm_index.rs:78-96; 0 = the reference's default (ratio 8, k 10 / 4)
  std::string input_file_src;
  uint64_t suffix_array_compression_ratio = 0;
  uint32_t lookup_table_kmer_len = 0;
  SymbolAlphabet alphabet = SymbolAlphabet::Nucleotide;
};

class FmIndex {
 public:
  // FmIndex::load (fm_index_file.rs:132)
  static FmIndex load(const std::string& path, const std::vector<int>& devices = {}) {
    awry_index* h = nullptr;
    check(awry_index_load(path.c_str(), devices.empty() ? nullptr : devices.data(), int(devices.size()), &h));
    return FmIndex(h);
  }
  // FmIndex::new (fm_index.rs:142-268) on the GPU; `save_to` non-empty = also FmIndex::save
  static FmIndex build(const FmBuildArgs& args, const std::vector<int>& devices = {}, const std::string& save_to = "") {
    awry_build_args a{};
    a.input_file_src = args.input_file_src.c_str();
    a.output_file_src = save_to.empty() ? nullptr : save_to.c_str();
    a.alphabet = uint32_t(args.alphabet);
    a.lookup_table_kmer_len = args.lookup_table_kmer_len;
    a.suffix_array_compression_ratio = args.suffix_array_compression_ratio;
    a.device = devices.empty() ? 0 : devices[0];
    awry_index* h = nullptr;
    check(awry_index_build(&a, devices.empty() ? nullptr : devices.data(), int(devices.size()), &h));
    return FmIndex(h);
  }
  // hand-over from the reference's CPU construction (FmIndex::new, fm_index.rs:142-268)
  static FmIndex from_parts(const awry_parts& parts, const std::vector<int>& devices = {}) {
    awry_index* h = nullptr;
    check(awry_index_from_parts(&parts, devices.empty() ? nullptr : devices.data(), int(devices.size()), &h));
    return FmIndex(h);
  }
  // FmIndex::save (fm_index_file.rs:42-106)
  void save(const std::string& path) const { check(awry_index_save(h_, path.c_str())); }
  FmIndex(FmIndex&& o) noexcept : h_(std::exchange(o.h_, nullptr)), info_(o.info_) {}
  FmIndex& operator=(FmIndex&& o) noexcept {
    if (this != &o) {
      awry_index_free(h_);
      h_ = std::exchange(o.h_, nullptr);
      info_ = o.info_;
    }
    return *this;
  }
  FmIndex(const FmIndex&) = delete;
  FmIndex& operator=(const FmIndex&) = delete;
  ~FmIndex() { awry_index_free(h_); }

  // getters (fm_index.rs:302-368)
  SymbolAlphabet alphabet() const { return SymbolAlphabet(info_.alphabet); }
  uint64_t suffix_array_compression_ratio() const { return info_.sa_ratio; }
  uint64_t bwt_len() const { return info_.bwt_len; }
  uint64_t version_number() const { return info_.version; }
  std::vector<uint64_t> prefix_sums() const {
    return std::vector<uint64_t>(info_.prefix_sums, info_.prefix_sums + info_.n_prefix_sums);
  }

  // fm_index.rs:383, :559-593
  SearchRange initial_search_range(char symbol) const {
    awry_range r;
    check(awry_initial_range(h_, uint8_t(symbol), &r));
    return {r.start_ptr, r.end_ptr};
  }
  SearchRange update_range_with_symbol(SearchRange range, char symbol) const {
    awry_range r;
    check(awry_update_range(h_, awry_range{range.start_ptr, range.end_ptr}, uint8_t(symbol), &r));
    return {r.start_ptr, r.end_ptr};
  }
  uint64_t backstep(uint64_t search_pointer) const {
    uint64_t out = 0;
    check(awry_backstep(h_, search_pointer, &out));
    return out;
  }

  // fm_index.rs:499-501, :516-544: batches of one
  uint64_t count_string(std::string_view query) const { return parallel_count({query})[0]; }
  std::vector<LocalizedSequencePosition> locate_string(std::string_view query) const {
    return parallel_locate({query})[0];
  }

  // fm_index.rs:455-460; input order preserved
  std::vector<uint64_t> parallel_count(const std::vector<std::string_view>& queries) const {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> off;
    pack(queries, bytes, off);
    std::vector<uint64_t> counts(queries.size());
    check(awry_count_batch(h_, bytes.data(), off.data(), queries.size(), counts.data()));
    return counts;
  }
  // fm_index.rs:479-487; per-query hits in BWT-row order (or sorted ascending)
  std::vector<std::vector<LocalizedSequencePosition>> parallel_locate(
      const std::vector<std::string_view>& queries, bool sorted = false) const {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> off;
    pack(queries, bytes, off);
    std::vector<uint64_t> hit_off(queries.size() + 1);
    awry_hit* hits = nullptr;
    uint64_t n = 0;
    check(awry_locate_batch(h_, bytes.data(), off.data(), queries.size(),
                            sorted ? AWRY_LOCATE_SORTED : AWRY_LOCATE_BWT_ORDER, hit_off.data(), &hits, &n));
    std::vector<std::vector<LocalizedSequencePosition>> out(queries.size());
    for (size_t q = 0; q < queries.size(); q++)
      for (uint64_t i = hit_off[q]; i < hit_off[q + 1]; i++) out[q].push_back({hits[i].seq_idx, hits[i].local_pos});
    awry_hits_free(hits);
    return out;
  }
  awry_index* handle() const { return h_; }

 private:
  explicit FmIndex(awry_index* h) : h_(h) { check(awry_index_info(h_, &info_)); }
  static void check(int rc) {
    if (rc != AWRY_OK) throw Error(rc, awry_last_error());
  }
  static void pack(const std::vector<std::string_view>& qs, std::vector<uint8_t>& bytes, std::vector<uint64_t>& off) {
    off.assign(1, 0);
    for (auto q : qs) {
      bytes.insert(bytes.end(), q.begin(), q.end());
      off.push_back(bytes.size());
    }
    if (bytes.empty()) bytes.push_back(0);
  }
  awry_index* h_ = nullptr;
  awry_info info_{};
};

}  // namespace awry
