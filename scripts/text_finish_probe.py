"""Count accelerator "finish in the text" at the headline workload (cfg2: 10 M x 150-bp reads vs 3.1 Gbp,
device-resident): search-kernel time with the comparison (awry_set_count_variant(0)) against backward search to
the last symbol (1), on exact reads and on reads of which 10 % carry one substitution; counts compared with each
other and, on a sample, with the oracle.  Also residency 4 / 6 / 8 blocks per SM for the new kernel."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
from oracle import pyoracle as po
n, nq, L, k = int(os.environ.get("PROBE_N", 3_100_000_000)), 10_000_000, 150, 13
t0 = time.time()
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=k)
print(f"built {n} rows in {time.time()-t0:.1f} s", flush=True)
t0 = time.time()
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
print(f"device replica in {time.time()-t0:.1f} s, device bytes {ix.device_bytes()}", flush=True)
off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * L
cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
f.profile_enable(True)
orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
QUICK = os.environ.get("PROBE_QUICK") == "1"  # branching kernel only, exact reads only (ticket-size sweeps)
for mut, what in ((0, "exact reads"), (100_000, "10 % of the reads with one substitution"))[:1 if QUICK else 2]:
    d = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(0, n, 3, nq, L, 4, d.data_ptr(), mut_ppm=mut)
    ref = None
    for variant, lanes, bps, name in ((1, 0, 0, "backward search only                      "),
                                      (0, 0, 0, "finish in the text, wave kernel           "),
                                      (0, 80, 0, "finish in the text, branching, 5 blk/SM   "),
                                      (0, 0, 6, "finish in the text, branching, 6 blk/SM   "),
                                      (0, 81, 0, "finish in the text, states, 1 slot, 8 blk "),
                                      (0, 83, 0, "finish in the text, states, 1 slot, 6 blk "),
                                      (0, 82, 0, "finish in the text, states, 2 slots, 5 blk"),
                                      (0, 84, 0, "finish in the text, states, 2 slots, 4 blk"))[:4 if QUICK else 8]:
        f.set_count_variant(variant)
        f.set_search_variant(lanes, 0, bps)
        for _ in range(3):
            ix.count_device(d.data_ptr(), off.data_ptr(), nq, cnt.data_ptr(), st)
        torch.cuda.synchronize()
        f.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ix.count_device(d.data_ptr(), off.data_ptr(), nq, cnt.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        p = f.profile_get()
        ms = p["search_ms"] / p["search_launches"]
        ref = cnt.clone() if ref is None else ref
        same = bool(torch.equal(cnt, ref))
        print(f"{what}: {name}: search kernel {ms:.2f} ms, whole call {e0.elapsed_time(e1)/10:.2f} ms per 10 M reads "
              f"= {nq/(e0.elapsed_time(e1)/10)/1e3:.0f} M reads/s; counts equal: {same}", flush=True)
    f.set_count_variant(0)
    f.set_search_variant(0, 0, 0)
    ns = 20_000
    want, _ = orc.count_batch(d[: ns * L].cpu().numpy(), np.arange(ns + 1, dtype=np.uint64) * np.uint64(L))
    print(f"{what}: parity vs oracle on {ns} reads: {bool(np.array_equal(want, ref[:ns].cpu().numpy().view(np.uint64)))}; "
          f"reads with count 0: {int((ref == 0).sum())}", flush=True)
    del d
