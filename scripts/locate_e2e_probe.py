"""cfg3 end to end (1 M x 50-bp queries, pinned buffers): sync-free pass 2 vs the two-sync pipeline, chunk sizes,
host packing on / off / pre-packed.  AWRY_B200_TRACE=1 prints the timeline of every call instead."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n = int(os.environ.get("PROBE_TEXT", 3_100_000_000))
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
nl, ll = 1_000_000, 50
d = torch.empty(nl * ll, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nl, ll, 5, d.data_ptr())
hq = torch.empty(nl * ll, dtype=torch.uint8, pin_memory=True); hq.copy_(d)
ho = torch.empty(nl + 1, dtype=torch.int64, pin_memory=True); ho.copy_(torch.arange(0, nl + 1, dtype=torch.int64) * ll)
qb, qo = hq.numpy(), ho.numpy().view(np.uint64)
hoff = torch.zeros(nl + 1, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
hits = torch.zeros((nl + 1024, 2), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
crumbs_np, exc = f.host_pack_dna(qb)
hc = torch.empty(len(crumbs_np) + 64, dtype=torch.uint8, pin_memory=True); hc[:len(crumbs_np)].copy_(torch.from_numpy(crumbs_np)); cr = hc.numpy()

def run(fn, reps=20):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3

if os.environ.get("AWRY_B200_TRACE"):
    for i in range(3):
        print(f"---- call {i}", file=sys.stderr)
        ix.locate_packed_into(qb, qo, hoff, hits)
    sys.exit(0)
ref = None
for direct in ("1", "0"):
    os.environ["AWRY_B200_LOCATE_DIRECT"] = direct
    for cq in (1 << 17, 1 << 18, 1 << 19, 1 << 20):
        os.environ["AWRY_B200_LOCATE_CHUNK_Q"] = str(cq)
        row = []
        for mode, name in ((1, "hostpack"), (0, "ascii")):
            f.set_host_pack(mode)
            row.append(f"{name} {run(lambda: ix.locate_packed_into(qb, qo, hoff, hits)):.3f} ms")
        f.set_host_pack(-1)
        row.append(f"prepacked {run(lambda: ix.locate_prepacked_into(cr, qo, exc, hoff, hits)):.3f} ms")
        if ref is None: ref = (hoff.copy(), hits.copy())
        assert np.array_equal(hoff, ref[0]) and np.array_equal(hits, ref[1])
        print(f"direct={direct} chunk_q={cq}: " + "; ".join(row), flush=True)
