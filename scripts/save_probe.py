"""awry_index_save at BASELINE scale (3.1 Gbp, k = 13, SA ratio 8; 4.56 GB `.awry` v1 file in /dev/shm):
from_parts -> save (A); the block, prefix-sum and SA sections of A must equal the reference-layout arrays the
handle was made from (the inverse re-layout at full size, 47 chunks); load(A) -> save (B) must reproduce A."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import FmIndex  # noqa: E402
from fixtures import pyfixture_gpu as fxg  # noqa: E402

n, k, ratio = 3_100_000_000, 13, 8
d = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
A, B = os.path.join(d, "awry_save_a.awry"), os.path.join(d, "awry_save_b.awry")
try:
    parts, _ = fxg.build_parts(0, n, 3, ratio=ratio, kmer_len=k)
    with FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums,
                            parts.sa_words) as ix:
        t0 = time.time()
        ix.save(A)
        t1 = time.time()
    size = os.path.getsize(A)
    print(f"save (handle from parts): {t1 - t0:.2f} s for {size / 1e9:.2f} GB = {size / 1e9 / (t1 - t0):.2f} GB/s", flush=True)
    m = np.memmap(A, dtype=np.uint8, mode="r")
    off = 43
    for name, arr in (("blocks", parts.blocks), ("prefix sums", parts.prefix_sums), ("SA words", parts.sa_words)):
        raw = arr.view(np.uint8)
        same = np.array_equal(m[off:off + raw.size], raw)
        print(f"  {name:12s} section @ {off}: {raw.size} bytes, equal to the source arrays: {same}", flush=True)
        assert same
        off += raw.size
    assert m[off] == k
    del m
    t2 = time.time()
    with FmIndex.load(A) as ix2:
        t3 = time.time()
        ix2.save(B)
        t4 = time.time()
    print(f"load: {t3 - t2:.2f} s; save (loaded handle): {t4 - t3:.2f} s", flush=True)
    a, b = np.memmap(A, dtype=np.uint8, mode="r"), np.memmap(B, dtype=np.uint8, mode="r")
    same = a.shape == b.shape and all(np.array_equal(a[i:i + (1 << 28)], b[i:i + (1 << 28)]) for i in range(0, a.size, 1 << 28))
    print(f"load -> save reproduces the file byte for byte: {same}", flush=True)
    assert same
finally:
    for p in (A, B):
        if os.path.exists(p):
            os.remove(p)
