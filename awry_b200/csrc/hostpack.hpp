// hostpack.hpp -- host-side 2-bit packing of nucleotide queries ahead of the PCIe copy (internal).
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

namespace awry {

// true when the CPU has the SIMD level the packer is written for (AVX2 + BMI2; AVX-512BW used if present)
bool host_pack_supported();
// worker threads of the process-wide pool (AWRY_B200_HOST_THREADS, default min(16, cores / LOCAL_WORLD_SIZE))
int host_pool_threads();
// runs fn(t, nt) on every pool thread (the caller is thread 0); returns when all are done
void host_parallel(const std::function<void(int, int)>& fn);

// src[0..n) ASCII -> dst: crumb i = (src[i] >> 1) & 3 at bits 2*(i%4) of dst[i/4]  (A0 C1 T2 G3, either
// case).  Every byte that is not one of ACGTacgt is reported as (i << 8) | byte in `exceptions` (in no
// particular order); its crumb is arbitrary and gets patched on the device.  Returns false (dst incomplete) when
// more than n / max_exc_div bytes are exceptions -- the caller then sends the chunk as ASCII.
bool host_pack_dna(const uint8_t* src, size_t n, uint8_t* dst, std::vector<uint64_t>& exceptions, size_t max_exc_div);

}  // namespace awry
