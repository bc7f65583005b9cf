// fixture_gpu.cu -- TEST/BENCH FIXTURES on the GPU. Not part of the product path.
//
// Device-side generators of the synthetic workloads of BASELINE.md: the seeded text (same counter-based
// generator as fixture_cpu.cpp, so CPU and GPU fixtures agree byte for byte) and exact-substring
// queries regenerated from the text's seed, written straight into device buffers.
// (The GPU index builder that used to live here is now the product's awry_build_parts; the tests
// check it bit for bit against the independent CPU builder in fixture_cpu.cpp.)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <string>

namespace {

thread_local char g_err[512];
int set_err(const std::string& m) {
  snprintf(g_err, sizeof g_err, "%s", m.c_str());
  return -1;
}

__host__ __device__ inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// must match fixture_cpu.cpp
__host__ __device__ inline uint64_t rnd_at(uint64_t seed, uint64_t i) {
  return mix64(mix64(seed) ^ (i * 0xD1342543DE82EF95ull));
}
// reference symbol index of synthetic text position i (alphabet.rs:169-248)
__host__ __device__ inline uint8_t synth_index(int alphabet, uint64_t seed, uint64_t i) {
  uint64_t r = rnd_at(seed, i);
  if (alphabet == 0) {
    uint32_t d = uint32_t(r >> 62);  // A C G T
    return d == 3 ? 5 : uint8_t(d + 1);
  }
  uint32_t d = uint32_t(((r >> 32) * 20) >> 32);  // ACDEFGHIKLMNPQRSTVWY -> 1..19, 21
  return d == 19 ? 21 : uint8_t(d + 1);
}
__host__ __device__ inline uint8_t index_to_ascii(int alphabet, uint8_t idx) {
  const char* D = "$ACGNT";
  const char* A = "$ACDEFGHIKLMNPQRSTVWXY";
  return uint8_t(alphabet == 0 ? D[idx] : A[idx]);
}

__global__ void gen_text_kernel(int alphabet, uint64_t seed, uint64_t n, uint8_t* out) {
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += stride)
    out[i] = index_to_ascii(alphabet, synth_index(alphabet, seed, i));
}

// queries regenerated from the text's seed: q = text[pos .. pos+qlen), optional single substitution
__global__ void gen_queries_kernel(int alphabet, uint64_t n, uint64_t text_seed, uint64_t nq, uint64_t qlen,
                                   uint64_t qseed, uint32_t mut_ppm, uint8_t* __restrict__ out) {
  uint64_t total = nq * qlen;
  uint64_t stride = gridDim.x * uint64_t(blockDim.x);
  uint64_t span = n - qlen + 1;
  for (uint64_t t = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; t < total; t += stride) {
    uint64_t q = t / qlen, k = t % qlen;
    uint64_t r = rnd_at(qseed, q);
    uint64_t pos = __umul64hi(r, span);
    uint8_t idx = synth_index(alphabet, text_seed, pos + k);
    if (mut_ppm) {
      uint64_t m = rnd_at(qseed ^ 0xA5A5A5A5ull, q);
      if (uint32_t(m % 1000000ull) < mut_ppm && (m >> 32) % qlen == k) {
        // a different encoding symbol at this position
        uint64_t alt = rnd_at(qseed ^ 0x5A5A5A5Aull, q);
        uint8_t nidx = idx;
        for (int tries = 0; nidx == idx; tries++) nidx = synth_index(alphabet, alt, uint64_t(tries));
        idx = nidx;
      }
    }
    out[t] = index_to_ascii(alphabet, idx);
  }
}


}  // namespace

extern "C" {

const char* fxg_last_error(void) { return g_err; }

// ASCII text of (alphabet, seed) into a DEVICE buffer of n bytes (== fx_gen_text on the host)
int fxg_gen_text_device(int alphabet, uint64_t n, uint64_t seed, void* d_out, void* stream) {
  gen_text_kernel<<<148 * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(alphabet, seed, n,
                                                                           static_cast<uint8_t*>(d_out));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(cudaGetErrorString(e));
  return 0;
}

// Synthetic queries straight into a DEVICE buffer (nq*qlen ASCII bytes): exact substrings of the
// synthetic text of (alphabet, n, text_seed) at seeded uniform positions; mut_ppm of them (per
// million) carry one substitution.  Same positions as fx_gen_substring_queries for equal seeds.
int fxg_gen_queries_device(int alphabet, uint64_t n, uint64_t text_seed, uint64_t nq, uint64_t qlen,
                           uint64_t qseed, uint32_t mut_ppm, void* d_out, void* stream) {
  if (qlen == 0 || qlen > n) return set_err("bad query length");
  gen_queries_kernel<<<148 * 16, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      alphabet, n, text_seed, nq, qlen, qseed, mut_ppm, static_cast<uint8_t*>(d_out));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err(cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
