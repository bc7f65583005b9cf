// hostpack.hpp -- host-side 2-bit packing of nucleotide queries ahead of the PCIe copy (internal).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace awry {

// true when the CPU has the SIMD level the packer is written for (AVX2 + BMI2; AVX-512BW used if present)
bool host_pack_supported();
// thread cap of the process-wide pool (AWRY_B200_HOST_THREADS, default min(16, cores / LOCAL_WORLD_SIZE);
// awry_set_host_threads changes it)
int host_pool_threads();
int host_default_threads();
void host_set_threads(int n);  // n <= 0: back to the default
// runs fn(block) for every block in [0, n_blocks) on the shared pool (the caller works too) and returns when
// all are done; several regions may run at once (one per replica thread).  An exception thrown by a block is
// rethrown here.  max_threads = 0: the pool's cap.
void host_parallel_blocks(size_t n_blocks, const std::function<void(size_t)>& fn, int max_threads = 0);

// src[0..n) ASCII -> dst: crumb i = (src[i] >> 1) & 3 at bits 2*(i%4) of dst[i/4]  (A0 C1 T2 G3, either
// case).  Every byte that is not one of ACGTacgt is reported as (i << 8) | byte in `exceptions`, ascending by
// position; its crumb is arbitrary and gets patched on the device.  Returns false (dst incomplete) when
// more than n / max_exc_div bytes are exceptions -- the caller then sends the chunk as ASCII.
// *sharers: parallel regions that were open in the pool when this one started, itself included (the share of
// the pool this call could count on was 1 / sharers).
bool host_pack_dna(const uint8_t* src, size_t n, uint8_t* dst, std::vector<uint64_t>& exceptions, size_t max_exc_div,
                   int max_threads = 0, int* sharers = nullptr);

}  // namespace awry
