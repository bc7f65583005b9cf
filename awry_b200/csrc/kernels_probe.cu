// kernels_probe.cu -- the random-gather roofline probe (awry_bench_random_gather): what a
// B200 delivers for independent, uniformly random, aligned reads -- the denominator of the search kernels.
#include <algorithm>
#include <cstdio>

#include "kernels_common.cuh"

namespace awry {

// ------------------------------------------------------------------ random-gather roofline probe

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// LANES consecutive lanes read one GRANULE-byte aligned granule at a uniformly random place;
// UNROLL independent granules per group are in flight before any is consumed.
template <int GRANULE, int LANES, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const char* __restrict__ buf, uint64_t n_granules,
                                                     uint64_t reads_per_group, uint64_t seed,
                                                     uint32_t* __restrict__ sink) {
  constexpr int BYTES = GRANULE / LANES;          // per lane
  constexpr int NV = BYTES >= 32 ? BYTES / 32 : 1;  // vector loads per lane
  const uint32_t lane = threadIdx.x & 31, sub = lane % LANES;
  const uint64_t group = (blockIdx.x * uint64_t(blockDim.x) + threadIdx.x) / LANES;
  uint32_t acc = 0;
  // cheap address stream (the probe must not be ALU-bound): 64-bit LCG per group, high product
  uint64_t x = mix64(seed ^ (group * 0x100000001B3ull));
  for (uint64_t it = 0; it < reads_per_group; it += UNROLL) {
    uint32_t v[UNROLL][NV * (BYTES >= 32 ? 8 : BYTES / 4)];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      x = x * 6364136223846793005ull + 1442695040888963407ull;
      uint64_t g = __umul64hi(x, n_granules);
      const char* p = buf + g * GRANULE + sub * BYTES;
      if (BYTES >= 32) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
          u32x8 r = ldg256(p + 32 * i);
#pragma unroll
          for (int k = 0; k < 8; k++) v[u][8 * i + k] = r.v[k];
        }
      } else if (BYTES == 16) {
        uint4 r = ldg128(reinterpret_cast<const uint4*>(p));
        v[u][0] = r.x, v[u][1] = r.y, v[u][2] = r.z, v[u][3] = r.w;
      } else {
        uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
        v[u][0] = r.x, v[u][1] = r.y;
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++)
#pragma unroll
      for (int k = 0; k < int(sizeof(v[0]) / 4); k++) acc ^= v[u][k];
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// ---- TMA variant of the probe: one THREAD per read; a 128-B bulk async copy (cp.async.bulk, the
// 1-D TMA path) lands the granule in the thread's shared-memory slot and completes a per-thread
// mbarrier, so a request in flight costs no registers.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_128(void* dst_smem, const void* src_gmem, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int DEPTH>
__global__ void __launch_bounds__(256) gather_tma_kernel(const char* __restrict__ buf, uint64_t n_granules,
                                                         uint64_t reads_per_thread, uint64_t seed,
                                                         uint32_t* __restrict__ sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* slots = smem;                                                   // [DEPTH][256][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DEPTH * 256 * 128);  // [DEPTH][256]
  const uint32_t t = threadIdx.x;
#pragma unroll
  for (int d = 0; d < DEPTH; d++) mbar_init(&bars[d * 256 + t], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();
  __syncthreads();
  uint64_t x = mix64(seed ^ ((blockIdx.x * uint64_t(blockDim.x) + t) * 0x100000001B3ull));
  uint32_t acc = 0;
  auto issue = [&](int d) {
    x = x * 6364136223846793005ull + 1442695040888963407ull;
    uint64_t g = __umul64hi(x, n_granules);
    mbar_expect_tx(&bars[d * 256 + t], 128);
    bulk_load_128(slots + (size_t(d) * 256 + t) * 128, buf + g * 128, &bars[d * 256 + t]);
  };
#pragma unroll
  for (int d = 0; d < DEPTH; d++) issue(d);
  uint32_t parity = 0;
  for (uint64_t it = 0; it < reads_per_thread; it += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; d++) {
      while (!mbar_try_wait(&bars[d * 256 + t], parity)) {
      }
      const uint4* s = reinterpret_cast<const uint4*>(slots + (size_t(d) * 256 + t) * 128);
      uint4 a = s[(t + d) & 7];  // one 16-B word of the landed granule
      acc ^= a.x ^ a.y ^ a.z ^ a.w;
      fence_proxy_async();  // order the generic read before the next async write to the slot
      if (it + DEPTH < reads_per_thread) issue(d);
    }
    parity ^= 1;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int DEPTH>
static cudaError_t gather_tma_run(const char* buf, uint64_t n_granules, uint64_t n_reads, int iters,
                                  uint32_t* sink, double* ms_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = size_t(DEPTH) * 256 * (128 + 8);
  cudaError_t e = cudaFuncSetAttribute(gather_tma_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gather_tma_kernel<DEPTH>, 256, smem);
  if (e != cudaSuccess) return e;
  unsigned grid = unsigned(sms * std::max(per_sm, 1));
  uint64_t threads = uint64_t(grid) * 256;
  uint64_t per_thread = ((n_reads + threads - 1) / threads + DEPTH - 1) / DEPTH * DEPTH;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int i = 0; i < iters + 1; i++) {
    cudaEventRecord(e0);
    gather_tma_kernel<DEPTH><<<grid, 256, smem>>>(buf, n_granules, per_thread, 0x5eed + i, sink);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return e;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i > 0 && ms < best) best = ms;
    COUNT_LAUNCH();
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  fprintf(stderr, "[gather_tma depth %d] %d blocks/SM, %llu reads/thread\n", DEPTH, per_sm, (unsigned long long)per_thread);
  *ms_out = double(best) / (double(per_thread * threads) / double(n_reads));
  return cudaGetLastError();
}

template <int GRANULE, int LANES, int UNROLL = 4>
static cudaError_t gather_run(const char* buf, uint64_t n_granules, uint64_t n_reads, int iters,
                              uint32_t* sink, double* ms_out, int blocks_per_sm = 8) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned grid = unsigned(sms) * unsigned(blocks_per_sm);
  uint64_t groups = uint64_t(grid) * 256 / LANES;
  uint64_t per_group = ((n_reads + groups - 1) / groups + UNROLL - 1) / UNROLL * UNROLL;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int i = 0; i < iters + 1; i++) {
    cudaEventRecord(e0);
    gather_kernel<GRANULE, LANES, UNROLL><<<grid, 256>>>(buf, n_granules, per_group, 0x5eed + i, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) return e;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (i > 0 && ms < best) best = ms;
    COUNT_LAUNCH();
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = double(best) / (double(per_group * groups) / double(n_reads));  // normalise to n_reads
  return cudaGetLastError();
}

cudaError_t run_random_gather(uint64_t footprint_bytes, uint32_t granule, uint32_t lanes,
                              uint64_t n_reads, int iters, double* reads_per_s, double* gb_per_s) {
  char* buf = nullptr;
  uint32_t* sink = nullptr;
  cudaError_t e = cudaMalloc(&buf, footprint_bytes);
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&sink, 4);
  if (e != cudaSuccess) {
    cudaFree(buf);
    return e;
  }
  cudaMemset(buf, 1, footprint_bytes);
  uint64_t n_granules = footprint_bytes / granule;
  double ms = 0;
  e = cudaErrorInvalidValue;
#define GATHER_CASE(G, L) \
  if (granule == G && lanes == L) e = gather_run<G, L>(buf, n_granules, n_reads, iters, sink, &ms);
  GATHER_CASE(32, 1) GATHER_CASE(32, 2) GATHER_CASE(32, 4)
  GATHER_CASE(64, 1) GATHER_CASE(64, 2) GATHER_CASE(64, 4) GATHER_CASE(64, 8)
  GATHER_CASE(128, 1) GATHER_CASE(128, 2) GATHER_CASE(128, 4) GATHER_CASE(128, 8)
#undef GATHER_CASE
  // lanes = 104 / 102: the 128-B / 4-lane probe with ONE / TWO reads in flight per lane group (the
  // dependent-chain shape of the search kernels) instead of four
  if (granule == 128 && lanes == 104) e = gather_run<128, 4, 1>(buf, n_granules, n_reads, iters, sink, &ms);
  if (granule == 128 && lanes == 102) e = gather_run<128, 4, 2>(buf, n_granules, n_reads, iters, sink, &ms);
  // lanes = 1000 + 10 * blocks_per_sm + unroll: 128-B / 4-lane probe at reduced residency (256-thread blocks)
  if (granule == 128 && lanes >= 1000 && lanes < 2000) {
    int bps = int(lanes - 1000) / 10, un = int(lanes - 1000) % 10;
    if (un == 1) e = gather_run<128, 4, 1>(buf, n_granules, n_reads, iters, sink, &ms, bps);
    if (un == 2) e = gather_run<128, 4, 2>(buf, n_granules, n_reads, iters, sink, &ms, bps);
    if (un == 4) e = gather_run<128, 4, 4>(buf, n_granules, n_reads, iters, sink, &ms, bps);
    if (un == 8) e = gather_run<128, 4, 8>(buf, n_granules, n_reads, iters, sink, &ms, bps);
  }
  // lanes = 3000 + 10 * waves + unroll: full residency, `waves` x as many blocks as fit at once, so the
  // hardware block scheduler balances the SMs dynamically (a static split ends with the slowest SM)
  if (granule == 128 && lanes >= 3000 && lanes < 4000) {
    int waves = int(lanes - 3000) / 10, un = int(lanes - 3000) % 10;
    if (un == 1) e = gather_run<128, 4, 1>(buf, n_granules, n_reads, iters, sink, &ms, 8 * waves);
    if (un == 2) e = gather_run<128, 4, 2>(buf, n_granules, n_reads, iters, sink, &ms, 8 * waves);
    if (un == 4) e = gather_run<128, 4, 4>(buf, n_granules, n_reads, iters, sink, &ms, 8 * waves);
    if (un == 8) e = gather_run<128, 4, 8>(buf, n_granules, n_reads, iters, sink, &ms, 8 * waves);
  }
  // lanes = 4000 + waves / 5000 + waves: the same with 2 lanes x 2 LDG.256 / 1 lane x 4 LDG.256 per read
  if (granule == 128 && lanes >= 4000 && lanes < 5000)
    e = gather_run<128, 2, 4>(buf, n_granules, n_reads, iters, sink, &ms, 8 * int(lanes - 4000));
  if (granule == 128 && lanes >= 5000 && lanes < 6000)
    e = gather_run<128, 1, 4>(buf, n_granules, n_reads, iters, sink, &ms, 8 * int(lanes - 5000));
  // lanes = 201 / 202: one thread per read through cp.async.bulk + mbarrier, 1 / 2 reads in flight per thread
  if (granule == 128 && lanes == 201) e = gather_tma_run<1>(buf, n_granules, n_reads, iters, sink, &ms);
  if (granule == 128 && lanes == 202) e = gather_tma_run<2>(buf, n_granules, n_reads, iters, sink, &ms);
  cudaFree(buf);
  cudaFree(sink);
  if (e != cudaSuccess) return e;
  *reads_per_s = double(n_reads) / (ms * 1e-3);
  *gb_per_s = *reads_per_s * granule * 1e-9;
  return cudaSuccess;
}

}  // namespace awry
