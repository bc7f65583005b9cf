"""Does a small H2D copy slow the search kernel (and vice versa)?  cfg3-sized chunk (512 k x 50 bp) on stream A,
10 MB pinned H2D on stream B: solo and concurrent, CUDA events."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n = 3_100_000_000
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
for nl in (1 << 19, 1 << 20):
    ll = 50
    d = torch.empty(nl * ll, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nl, ll, 5, d.data_ptr())
    off = torch.arange(0, nl + 1, dtype=torch.int64, device="cuda") * ll
    cnt = torch.zeros(nl, dtype=torch.int64, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for mb in (10, 40):
        h = torch.empty(mb << 20, dtype=torch.uint8, pin_memory=True); dd = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
        def kern():
            ix.count_device(d.data_ptr(), off.data_ptr(), nl, cnt.data_ptr(), sa.cuda_stream)
        def copy():
            with torch.cuda.stream(sb):
                dd.copy_(h, non_blocking=True)
        def timed(fa, fb):
            torch.cuda.synchronize()
            ea0, ea1, eb0, eb1 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if fa: ea0.record(sa)
            if fb: eb0.record(sb)
            if fa: fa(); ea1.record(sa)
            if fb: fb(); eb1.record(sb)
            torch.cuda.synchronize()
            return (ea0.elapsed_time(ea1) if fa else 0, eb0.elapsed_time(eb1) if fb else 0)
        for _ in range(3): timed(kern, copy)
        solo_k = min(timed(kern, None)[0] for _ in range(5))
        solo_c = min(timed(None, copy)[1] for _ in range(5))
        both = [timed(kern, copy) for _ in range(5)]
        print(f"nq={nl} copy={mb}MB: kernel solo {solo_k:.3f} ms, copy solo {solo_c:.3f} ms; concurrent kernel {min(b[0] for b in both):.3f} ms, copy {min(b[1] for b in both):.3f} ms", flush=True)
