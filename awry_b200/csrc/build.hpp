// build.hpp -- GPU index construction (internal to libawry_b200).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace awry {

// Reference-layout parts of text[0..n) + '$' (ASCII; host or device pointer): blocks
// (ceil((n+1)/256) x 20|44 u64), prefix_sums (7|23), sa_words.  phase_s: 8 doubles or nullptr
// (ingest, keys, sort, ties, bwt, milestones, sa-pack+copy, total).  0 on success, else err is set.
int build_parts(int alphabet, const uint8_t* text, uint64_t n, uint64_t ratio, int device, uint64_t* blocks_out,
                uint64_t* prefix_sums_out, uint64_t* sa_words_out, double* phase_s, std::string& err);

int read_sequence_file(const std::string& path, char delimiter, std::string& text, std::vector<uint64_t>& starts,
                       std::vector<std::string>& headers, std::string& err);

}  // namespace awry
