"""BASELINE cfg 5 at scale: repeat-rich synthetic DNA, SA ratio 32, 1 M x 50-bp queries with heavily
skewed hit counts.  Index built by the GPU fixture builder (prefix doubling), count + locate on the
device, parity of a sample against the CPU oracle."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from awry_b200 import FmIndex, fm_index as f  # noqa: E402
from fixtures import pyfixture_gpu as fxg, repeats  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--nq", type=int, default=1_000_000)
    ap.add_argument("--qlen", type=int, default=50)
    ap.add_argument("--ratio", type=int, default=32)
    ap.add_argument("--check", type=int, default=2000)
    a = ap.parse_args()
    t0 = time.time()
    scale = a.n / 1e9
    fams = ((300, int(100_000 * scale) + 10, 0.15), (6000, int(20_000 * scale) + 5, 0.15))
    text, regions = repeats.repeat_rich_text(a.n, seed=8, tandem_arrays=max(4, int(40 * scale)), families=fams)
    t1 = time.time()
    parts, phases = fxg.build_parts(0, a.n, 0, ratio=a.ratio, kmer_len=13, host_text=text)
    t2 = time.time()
    print(f"text {t1-t0:.1f}s; GPU fixture build {t2-t1:.1f}s phases {json.dumps({k: round(v, 2) for k, v in phases.items()})}", flush=True)
    qb, qo = repeats.repeat_queries(text, regions, a.nq, a.qlen, seed=9)
    ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words)
    d_q = torch.from_numpy(qb).cuda()
    d_off = torch.from_numpy(qo.astype(np.int64)).cuda()
    d_cnt = torch.zeros(a.nq, dtype=torch.int64, device="cuda")
    d_hoff = torch.zeros(a.nq + 1, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f.profile_enable(True)
    for it in range(3):
        f.profile_reset()
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), a.nq, d_cnt.data_ptr(), st)
        torch.cuda.synchronize()
        pc = f.profile_get()
    cnt = d_cnt.cpu().numpy()
    print(f"count: search {pc['search_ms']:.2f} ms; hits/query min {cnt.min()} median {int(np.median(cnt))} "
          f"mean {cnt.mean():.1f} max {cnt.max()} total {cnt.sum()}", flush=True)
    print(f"device bytes {ix.device_bytes()}", flush=True)
    for variant in (1, 0):          # LF-walk, then the unsampled-SA gather (kept for the parity check)
        f.set_locate_variant(variant)
        for it in range(3):
            f.profile_reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ptr, nh = ix.locate_device(d_q.data_ptr(), d_off.data_ptr(), a.nq, d_hoff.data_ptr(), stream=st)
            e1.record()
            torch.cuda.synchronize()
            pl = f.profile_get()
            ms = e0.elapsed_time(e1)
            if it < 2 or variant == 1:
                ix.device_free(ptr)
        print(f"locate variant {variant} ({'LF-walk' if variant else 'unsampled-SA gather'}): {nh} hits in {ms:.2f} ms "
              f"({nh/ms/1e3:.1f} M hits/s); search {pl['search_ms']:.2f} ms, pass 2 {pl['walk_ms']:.2f} ms "
              f"({nh/pl['walk_ms']/1e3:.1f} M hits/s in the pass-2 kernel)", flush=True)
    orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                    parts.prefix_sums, parts.sa_words)
    ns = a.check
    woff, whits, stt = orc.locate_batch(qb[: ns * a.qlen], qo[: ns + 1])
    hoff = d_hoff.cpu().numpy().astype(np.uint64)
    n_s = int(hoff[ns])
    buf = torch.empty(n_s * 2, dtype=torch.int64, device="cuda")
    C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), C.c_size_t(n_s * 16), 3)
    got = buf.cpu().numpy().astype(np.uint64).reshape(-1, 2)
    ok = np.array_equal(hoff[: ns + 1], woff) and np.array_equal(got, whits)
    print(f"parity on {ns} queries ({n_s} hits, mean walk {stt['walk_steps']/max(1,stt['hits']):.1f}): {'OK' if ok else 'MISMATCH'}")
    ix.device_free(ptr)
    print(f"total {time.time()-t0:.1f}s")


if __name__ == "__main__":
    main()
