"""One awry_count_batch over N replicas from pinned ASCII, for fixed packed shares of the bytes
(AWRY_B200_PACK_SHARE is read when a handle is created): where is the optimum on this host?"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
N = torch.cuda.device_count()
n, nq, L = 3_100_000_000, 10_000_000, 150
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
total = N * nq
h_q = torch.empty(total * L, dtype=torch.uint8, pin_memory=True)
for r in range(N):
    d = torch.empty(nq * L, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nq, L, 4 + 1000 * r, d.data_ptr())
    h_q[r * nq * L:(r + 1) * nq * L].copy_(d); del d
h_off = torch.empty(total + 1, dtype=torch.int64, pin_memory=True); h_off.copy_(torch.arange(0, total + 1, dtype=torch.int64) * L)
h_cnt = torch.empty(total, dtype=torch.int64, pin_memory=True)
torch.cuda.synchronize()
qb, qo, out = h_q.numpy(), h_off.numpy().view(np.uint64), h_cnt.numpy().view(np.uint64)
f.set_host_threads(min(os.cpu_count(), 256))
print(f"{N} GPUs, {os.cpu_count()} host threads", flush=True)
os.environ["AWRY_B200_FULL_SA"] = "0"; os.environ["AWRY_B200_LEAN_SA"] = "0"
ref = None
for share in ("auto", "1.0", "0.75", "0.5", "0.25", "0.0"):
    if share == "auto":
        os.environ.pop("AWRY_B200_PACK_SHARE", None)
    else:
        os.environ["AWRY_B200_PACK_SHARE"] = share
    ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words, devices=list(range(N)))
    for _ in range(4):
        ix.count_packed(qb, qo, out=out)
    f.profile_reset()
    t0 = time.perf_counter()
    for _ in range(8):
        ix.count_packed(qb, qo, out=out)
    dt = (time.perf_counter() - t0) / 8
    p = f.profile_get()
    s = int(out.sum(dtype=np.uint64)); ref = s if ref is None else ref
    assert s == ref
    print(f"packed share {share}: {dt*1e3:.1f} ms/step = {total/dt/1e6:.0f} M reads/s, h2d {p['h2d_bytes']/8/1e6:.0f} MB/step", flush=True)
    ix.close()
