"""cfg2 end to end from pinned ASCII: time per call and (AWRY_B200_TRACE=1) the host / device timeline of one call."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n, nq, L = 3_100_000_000, 10_000_000, 150
os.environ.setdefault("AWRY_B200_LEAN_SA", "0")
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
d = torch.empty(nq * L, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nq, L, 4, d.data_ptr())
hq = torch.empty(nq * L, dtype=torch.uint8, pin_memory=True); hq.copy_(d)
ho = torch.empty(nq + 1, dtype=torch.int64, pin_memory=True); ho.copy_(torch.arange(0, nq + 1, dtype=torch.int64) * L)
hc = torch.empty(nq, dtype=torch.int64, pin_memory=True)
qb, qo, out = hq.numpy(), ho.numpy().view(np.uint64), hc.numpy().view(np.uint64)
reps = 3 if os.environ.get("AWRY_B200_TRACE") else 12
for i in range(reps):
    if os.environ.get("AWRY_B200_TRACE"): print(f"---- call {i}", file=sys.stderr)
    t0 = time.perf_counter(); ix.count_packed(qb, qo, out=out); dt = time.perf_counter() - t0
    print(f"call {i}: {dt*1e3:.2f} ms", file=sys.stderr)
