// reads.cu -- device-side FASTQ / FASTA parser of the streaming query front-end
// ("a reader that feeds parallel_count / parallel_locate straight from FASTQ
// with overlapped parsing, packing and H2D").
//
// The reference has no such reader: its users parse reads on the CPU and hand `&str`s to
// FmIndex::parallel_count (/root/reference/src/fm_index.rs:455-460).  Here the raw file bytes go to
// the device in large chunks and are parsed there:
//   1. positions of '\n'                    cub::DeviceSelect::If over a counting iterator
//   2. one thread per line: kind + length   FASTQ: line j is header / sequence / '+' / quality by j % 4;
//                                           FASTA: '>' starts a header, every other line is sequence
//   3. two exclusive sums over the lines    compacted sequence offset of each line, record number
//   4. plan (one thread): complete records in the chunk, bytes consumed, sequence bytes
//   5. one warp per line                    copies sequence bytes to a contiguous query buffer and
//                                           writes the CSR offsets the search kernels take
// A record cut by the chunk boundary is carried over by the host (reads_api.cu).
#include <algorithm>
#include <cstdint>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cuda_runtime.h>

#include "reads.hpp"

namespace awry {

namespace {

struct IsNewline {
  const uint8_t* raw;
  __host__ __device__ bool operator()(uint32_t i) const { return raw[i] == '\n'; }
};

// line j = raw[start_j, end_j) with start_j = nl[j-1] + 1 (0 for j = 0), end_j = nl[j], a trailing '\r' dropped
__device__ __forceinline__ void line_span(const uint8_t* raw, const uint32_t* nl, uint32_t j, uint32_t& b, uint32_t& e) {
  b = j ? nl[j - 1] + 1 : 0;
  e = nl[j];
  if (e > b && raw[e - 1] == '\r') e--;
}

__global__ void classify_lines_kernel(const uint8_t* __restrict__ raw, const uint32_t* __restrict__ nl,
                                      const uint32_t* __restrict__ n_lines_p, int fastq,
                                      uint32_t* __restrict__ seq_len, uint32_t* __restrict__ is_hdr) {
  const uint32_t n = *n_lines_p;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    uint32_t b, e;
    line_span(raw, nl, j, b, e);
    bool hdr, seq;
    if (fastq) {
      hdr = (j & 3) == 0;
      seq = (j & 3) == 1;
    } else {
      hdr = e > b && raw[b] == '>';
      seq = !hdr;
    }
    is_hdr[j] = hdr ? 1u : 0u;
    seq_len[j] = seq ? e - b : 0u;
  }
}

// seq_off / hdr_rank are exclusive sums with n_lines + 1 entries (the last one is the total)
__global__ void plan_kernel(const uint32_t* __restrict__ nl, const uint32_t* __restrict__ n_lines_p,
                            const uint32_t* __restrict__ is_hdr, const uint32_t* __restrict__ hdr_rank,
                            const uint64_t* __restrict__ seq_off, int fastq, int at_eof, uint32_t n_bytes,
                            ReadsPlan* __restrict__ plan) {
  const uint32_t n = *n_lines_p;
  ReadsPlan p{};
  p.n_lines = n;
  if (fastq) {
    p.n_records = n / 4;
    p.first_unused_line = 4 * p.n_records;
    // at EOF a truncated record (1-3 trailing lines) is an error the host reports
    p.trailing_lines = n - p.first_unused_line;
  } else {
    uint32_t total_hdr = n ? hdr_rank[n] : 0;
    if (at_eof || total_hdr == 0) {
      p.n_records = total_hdr;
      p.first_unused_line = n;
    } else {
      // the last record may continue in the next chunk: find its header line (scan back)
      uint32_t j = n;
      while (j > 0 && !is_hdr[j - 1]) j--;
      p.n_records = total_hdr - 1;
      p.first_unused_line = j - 1;
    }
    p.trailing_lines = n - p.first_unused_line;
    // sequence lines before the first header of the very first chunk belong to no record
    p.orphan_bases = 0;
    if (n) {
      uint32_t j = 0;
      while (j < n && !is_hdr[j]) j++;
      p.orphan_bases = seq_off[j];
    }
  }
  p.consumed = p.first_unused_line ? nl[p.first_unused_line - 1] + 1 : 0;
  p.seq_bytes = (n ? seq_off[p.first_unused_line] : 0) - p.orphan_bases;
  p.n_bytes = n_bytes;
  *plan = p;
}

// one warp per line: header lines write the record's CSR offset, sequence lines copy their bytes
__global__ void __launch_bounds__(256)
    emit_reads_kernel(const uint8_t* __restrict__ raw, const uint32_t* __restrict__ nl, const uint32_t* __restrict__ is_hdr,
                      const uint32_t* __restrict__ hdr_rank, const uint64_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_len,
                      const ReadsPlan* __restrict__ plan_p, int fastq, uint8_t* __restrict__ qbytes,
                      uint64_t* __restrict__ qoff) {
  const ReadsPlan plan = *plan_p;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint64_t base = plan.orphan_bases;
  for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j <= plan.first_unused_line; j += nwarps) {
    if (j == plan.first_unused_line) {  // closing offset
      if (lane == 0) qoff[plan.n_records] = plan.seq_bytes;
      continue;
    }
    if (is_hdr[j]) {
      if (lane == 0) qoff[hdr_rank[j]] = seq_off[j] - base;
      continue;
    }
    uint32_t len = seq_len[j];
    if (len == 0) continue;
    if (!fastq && hdr_rank[j] == 0) continue;  // orphan sequence line before the first header
    uint32_t b, e;
    line_span(raw, nl, j, b, e);
    uint8_t* dst = qbytes + (seq_off[j] - base);
    const uint8_t* src = raw + b;
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
  }
}

}  // namespace

size_t reads_temp_bytes(uint32_t max_bytes) {
  size_t a = 0, b = 0, c = 0;
  thrust::counting_iterator<uint32_t> it(0);
  cub::DeviceSelect::If(nullptr, a, it, static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(max_bytes),
                        IsNewline{nullptr});
  cub::DeviceScan::ExclusiveSum(nullptr, b, static_cast<uint32_t*>(nullptr), static_cast<uint64_t*>(nullptr), int(max_bytes));
  cub::DeviceScan::ExclusiveSum(nullptr, c, static_cast<uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), int(max_bytes));
  return std::max(a, std::max(b, c)) + 256;
}

// Stage A: newline positions.  d_n_lines receives the count (device); the host reads it back to size stage B.
cudaError_t reads_find_lines(const uint8_t* d_raw, uint32_t n_bytes, uint32_t* d_nl, uint32_t* d_n_lines, void* d_temp,
                             size_t temp_bytes, cudaStream_t s) {
  thrust::counting_iterator<uint32_t> it(0);
  return cub::DeviceSelect::If(d_temp, temp_bytes, it, d_nl, d_n_lines, int(n_bytes), IsNewline{d_raw}, s);
}

// Stage B: everything after the line count is known on the host.
cudaError_t reads_parse_lines(const uint8_t* d_raw, uint32_t n_bytes, const uint32_t* d_nl, const uint32_t* d_n_lines,
                              uint32_t n_lines, int fastq, int at_eof, uint32_t* d_seq_len, uint32_t* d_is_hdr,
                              uint64_t* d_seq_off, uint32_t* d_hdr_rank, ReadsPlan* d_plan, uint8_t* d_qbytes,
                              uint64_t* d_qoff, void* d_temp, size_t temp_bytes, cudaStream_t s) {
  const unsigned grid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(n_lines) + 255) / 256, 148 * 16)));
  classify_lines_kernel<<<grid, 256, 0, s>>>(d_raw, d_nl, d_n_lines, fastq, d_seq_len, d_is_hdr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // n_lines + 1 outputs: the inputs carry one zero past the end (buffers are zero-initialised slots)
  e = cudaMemsetAsync(d_seq_len + n_lines, 0, 4, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(d_is_hdr + n_lines, 0, 4, s);
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_seq_len, d_seq_off, int(n_lines + 1), s);
  if (e != cudaSuccess) return e;
  e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_is_hdr, d_hdr_rank, int(n_lines + 1), s);
  if (e != cudaSuccess) return e;
  plan_kernel<<<1, 1, 0, s>>>(d_nl, d_n_lines, d_is_hdr, d_hdr_rank, d_seq_off, fastq, at_eof, n_bytes, d_plan);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const unsigned wgrid = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((uint64_t(n_lines) + 1 + 7) / 8, 148 * 32)));
  emit_reads_kernel<<<wgrid, 256, 0, s>>>(d_raw, d_nl, d_is_hdr, d_hdr_rank, d_seq_off, d_seq_len, d_plan, fastq, d_qbytes,
                                          d_qoff);
  return cudaGetLastError();
}

}  // namespace awry
