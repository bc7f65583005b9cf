// build.hpp -- GPU index construction (internal to libawry_b200).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace awry {

// Results of a construction left on the device (reference layout), for a hand-over without a host round trip
struct DeviceParts {
  uint64_t* d_blocks = nullptr;    // ceil((n+1)/256) x 20|44 u64
  uint64_t* d_sa_words = nullptr;  // bit-packed sampled suffix array
  uint64_t n_block_words = 0, n_sa_words = 0;
  int device = 0;
  void release();
};

// Reference-layout parts of text[0..n) + '$' (ASCII; host or device pointer): blocks
// (ceil((n+1)/256) x 20|44 u64), prefix_sums (7|23), sa_words.  blocks_out / sa_words_out may be nullptr
// (no copy to the host); `keep`, when given, receives the device arrays instead of their being freed.
// phase_s: 8 doubles or nullptr (ingest, keys, sort, ties, bwt, milestones, sa-pack+copy, total).
// 0 on success, else err is set.
int build_parts(int alphabet, const uint8_t* text, uint64_t n, uint64_t ratio, int device, uint64_t* blocks_out,
                uint64_t* prefix_sums_out, uint64_t* sa_words_out, double* phase_s, std::string& err,
                DeviceParts* keep = nullptr);

// the concatenated text of a sequence file (not zero-initialised: 3 GB of memset would cost more than the parse)
struct TextBuf {
  std::unique_ptr<uint8_t[]> p;
  size_t n = 0;
  const uint8_t* data() const { return p.get(); }
  size_t size() const { return n; }
  void reset() {
    p.reset();
    n = 0;
  }
};

int read_sequence_file(const std::string& path, char delimiter, TextBuf& text, std::vector<uint64_t>& starts,
                       std::vector<std::string>& headers, std::string& err);

}  // namespace awry
