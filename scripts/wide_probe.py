"""A wide index (4.6 G rows > 2^32: 64-bit row pointers) at the headline workload: 10 M x 150-bp reads, device-resident,
cooperative kernel (4 lanes per query) against the one-thread kernel; parity between the two and vs the oracle on a sample."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
from oracle import pyoracle as po
n, nq, L, k = 4_600_000_000, 10_000_000, 150, 13
t0 = time.time()
parts, phases = fxg.build_parts(0, n, 12, ratio=16, kmer_len=k)
print(f"built {n} rows in {time.time()-t0:.1f} s (phases {dict((a, round(b, 2)) for a, b in phases.items())})", flush=True)
t0 = time.time()
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
print(f"device replica in {time.time()-t0:.1f} s, row pointer bits {ix.row_pointer_bits()}, device bytes {ix.device_bytes()}", flush=True)
d = torch.empty(nq * L, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 12, nq, L, 4, d.data_ptr())
off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * L
cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ref = None
f.profile_enable(True)
for variant, name in ((0, "cooperative, 4 lanes per query"), (-1, "one thread per query")):
    f.set_search_variant(variant)
    for _ in range(2):
        ix.count_device(d.data_ptr(), off.data_ptr(), nq, cnt.data_ptr(), st)
    f.profile_reset()
    reps = 5 if variant == 0 else 2
    for _ in range(reps):
        ix.count_device(d.data_ptr(), off.data_ptr(), nq, cnt.data_ptr(), st)
    torch.cuda.synchronize()
    p = f.profile_get()
    ms = p["search_ms"] / p["search_launches"]
    ref = cnt.clone() if ref is None else ref
    assert torch.equal(cnt, ref)
    print(f"{name}: search kernel {ms:.2f} ms per 10 M reads = {nq/ms/1e3:.0f} M reads/s = {nq*(L-k)/ms/1e6:.1f} G LF steps/s", flush=True)
f.set_search_variant(0)
ns = 20_000
orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
want, _ = orc.count_batch(d[: ns * L].cpu().numpy(), np.arange(ns + 1, dtype=np.uint64) * np.uint64(L))
print("parity vs oracle on", ns, "reads:", bool(np.array_equal(want, ref[:ns].cpu().numpy().view(np.uint64))), "; min count", int(ref.min()))
