// reads.hpp -- device-side FASTQ / FASTA parsing for the streaming query front-end (internal).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace awry {

struct ReadsPlan {
  uint32_t n_lines;            // complete lines ('\n'-terminated) in the chunk
  uint32_t n_records;          // complete records emitted
  uint32_t first_unused_line;  // lines from here on belong to a record the chunk does not finish
  uint32_t trailing_lines;
  uint32_t consumed;           // bytes of the chunk covered by the emitted records
  uint32_t n_bytes;
  uint64_t seq_bytes;          // total sequence bytes emitted (== qoff[n_records])
  uint64_t orphan_bases;       // FASTA: sequence bytes before the first header (ignored)
};

size_t reads_temp_bytes(uint32_t max_bytes);
cudaError_t reads_find_lines(const uint8_t* d_raw, uint32_t n_bytes, uint32_t* d_nl, uint32_t* d_n_lines, void* d_temp,
                             size_t temp_bytes, cudaStream_t s);
cudaError_t reads_parse_lines(const uint8_t* d_raw, uint32_t n_bytes, const uint32_t* d_nl, const uint32_t* d_n_lines,
                              uint32_t n_lines, int fastq, int at_eof, uint32_t* d_seq_len, uint32_t* d_is_hdr,
                              uint64_t* d_seq_off, uint32_t* d_hdr_rank, ReadsPlan* d_plan, uint8_t* d_qbytes,
                              uint64_t* d_qoff, void* d_temp, size_t temp_bytes, cudaStream_t s);

}  // namespace awry
