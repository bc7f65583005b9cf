"""Rate of the library's host packer (awry_host_pack_dna) on this host: AWRY_B200_HOST_SIMD=0/1/2, AWRY_B200_HOST_THREADS"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000_000
L = f.native()
src = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, n, dtype=np.uint8)]
dst = np.zeros(n // 4 + 64, dtype=np.uint8)
ne = C.c_uint64()
chunk = 128 << 20
best = 1e9
for it in range(4):
    t0 = time.perf_counter()
    for o in range(0, n, chunk):       # the pipeline packs chunk by chunk
        m = min(chunk, n - o)
        L.awry_host_pack_dna(src.ctypes.data + o, m, dst.ctypes.data + o // 4, None, 0, C.byref(ne))
    best = min(best, time.perf_counter() - t0)
print(f"SIMD={os.environ.get('AWRY_B200_HOST_SIMD', 'auto')} threads={os.environ.get('AWRY_B200_HOST_THREADS', 'auto')}: "
      f"{best*1e3:.1f} ms for {n/1e9:.2f} GB = {n/best/1e9:.1f} GB/s", flush=True)
