"""parallel_locate end to end (pinned host buffers in and out, cfg3: 1 M x 50-bp queries vs 3.1 Gbp) against the
pipeline chunk size (AWRY_B200_LOCATE_CHUNK_Q is read per call); results must not depend on it."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex
from fixtures import pyfixture_gpu as fxg
n = 3_100_000_000
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
nl, ll = 1_000_000, 50
d = torch.empty(nl * ll, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nl, ll, 5, d.data_ptr())
hq = torch.empty(nl * ll, dtype=torch.uint8, pin_memory=True); hq.copy_(d)
ho = torch.empty(nl + 1, dtype=torch.int64, pin_memory=True); ho.copy_(torch.arange(0, nl + 1, dtype=torch.int64) * ll)
qb, qo = hq.numpy(), ho.numpy().view(np.uint64)
hoff = torch.zeros(nl + 1, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
hits = torch.zeros((nl + 1024, 2), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
ref = None
for chunk_q in [1 << 18, 3 << 17, 1 << 19, 1 << 20]:
    os.environ["AWRY_B200_LOCATE_CHUNK_Q"] = str(chunk_q)
    os.environ["AWRY_B200_LOCATE_CHUNK_MB"] = "64"
    ts = []
    for i in range(8):
        t0 = time.perf_counter(); n_hits = ix.locate_packed_into(qb, qo, hoff, hits); ts.append((time.perf_counter() - t0) * 1e3)
    got = (hoff.copy(), hits[:int(hoff[-1])].copy())
    if ref is None:
        ref = got
    same = np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    print(f"chunk_q {chunk_q:8d}: min {min(ts[2:]):.2f} ms  median {sorted(ts[2:])[3]:.2f} ms  ({int(hoff[-1])} hits; same results: {same})", flush=True)
