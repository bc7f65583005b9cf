// Reads like the reference's own test (fm_index.rs:612-664): every k-mer of a text is counted
// and located through the C++ host mirror and compared with brute-force positions.
// usage: cxx_host_mirror <index.awry> <text-file> <kmer-len>
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>

#include "awry_b200.hpp"

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  try {
    awry::FmIndex fm = awry::FmIndex::load(argv[1]);
    std::ifstream in(argv[2], std::ios::binary);
    std::stringstream ss;
    ss << in.rdbuf();
    std::string text = ss.str();
    size_t k = size_t(std::atoi(argv[3]));
    std::map<std::string, std::vector<uint64_t>> kmers;
    for (size_t p = 0; p + k <= text.size(); p++) kmers[text.substr(p, k)].push_back(p);
    std::vector<std::string_view> qs;
    for (auto& kv : kmers) qs.emplace_back(kv.first);
    auto counts = fm.parallel_count(qs);
    auto locs = fm.parallel_locate(qs);
    size_t i = 0;
    for (auto& kv : kmers) {
      if (counts[i] != kv.second.size()) return std::printf("count mismatch for %s\n", kv.first.c_str()), 1;
      std::vector<uint64_t> got;
      for (auto& h : locs[i]) got.push_back(h.local_position());
      std::sort(got.begin(), got.end());
      if (got != kv.second) return std::printf("locate mismatch for %s\n", kv.first.c_str()), 1;
      i++;
    }
    if (fm.count_string(qs[0]) != counts[0]) return 1;
    bool threw = false;
    try {
      fm.count_string("");
    } catch (const awry::Error& e) {
      threw = e.code == AWRY_ERR_INVALID_QUERY;
    }
    if (!threw) return std::printf("empty query did not raise\n"), 1;
    auto r = fm.initial_search_range('A');
    auto r2 = fm.update_range_with_symbol(r, 'C');
    std::printf("ok %zu kmers, range(A)=[%llu,%llu] CA len %llu\n", kmers.size(), (unsigned long long)r.start_ptr,
                (unsigned long long)r.end_ptr, (unsigned long long)r2.len());
    return 0;
  } catch (const awry::Error& e) {
    std::printf("awry error %d: %s\n", e.code, e.what());
    return 3;
  }
}
