"""Bench-scale probe: build the synthetic index on the GPU, time the search-kernel variants on
device-resident reads, check a sample against the CPU oracle."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from awry_b200 import FmIndex, fm_index as f  # noqa: E402
from fixtures import pyfixture_gpu as fxg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3_100_000_000)
    ap.add_argument("--alphabet", type=int, default=0)
    ap.add_argument("--nq", type=int, default=10_000_000)
    ap.add_argument("--qlen", type=int, default=150)
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--ratio", type=int, default=8)
    ap.add_argument("--variants", default="8,2")
    ap.add_argument("--bps", default="0")
    ap.add_argument("--check", type=int, default=20000)
    ap.add_argument("--locate-nq", type=int, default=10_000_000)
    ap.add_argument("--locate-qlen", type=int, default=50)
    ap.add_argument("--via-file", default="", help="also write the index as an .awry v1 file here and load it back")
    a = ap.parse_args()
    t0 = time.time()
    parts, phases = fxg.build_parts(a.alphabet, a.n, 3, ratio=a.ratio, kmer_len=a.k)
    print("fixture build phases (s):", json.dumps({k: round(v, 2) for k, v in phases.items()}), flush=True)
    for full_sa in ("0", None):     # cost of rebuilding the unsampled suffix array at load time
        if full_sa is None:
            os.environ.pop("AWRY_B200_FULL_SA", None)
        else:
            os.environ["AWRY_B200_FULL_SA"] = full_sa
        t1 = time.time()
        ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                parts.prefix_sums, parts.sa_words)
        print(f"from_parts (AWRY_B200_FULL_SA={full_sa}) {time.time()-t1:.2f}s device bytes {ix.device_bytes()}", flush=True)
        if full_sa is not None:
            ix.close()
    if a.via_file:
        t2 = time.time()
        parts.write(a.via_file, reference_table=False)
        sz = os.path.getsize(a.via_file)
        t3 = time.time()
        ix2 = FmIndex.load(a.via_file)
        t4 = time.time()
        print(f"file: wrote {sz/1e9:.2f} GB in {t3-t2:.1f}s; awry_index_load {t4-t3:.2f}s = {sz/(t4-t3)/1e9:.2f} GB/s "
              f"(incl. re-layout, seed table, pair index)", flush=True)
        ix.close()
        ix = ix2
        os.remove(a.via_file)
    d_q = torch.empty(a.nq * a.qlen, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(a.alphabet, a.n, 3, a.nq, a.qlen, 4, d_q.data_ptr())
    d_off = torch.arange(0, a.nq + 1, dtype=torch.int64, device="cuda") * a.qlen
    d_cnt = torch.zeros(a.nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f.profile_enable(True)
    results = []
    for lanes in [int(x) for x in a.variants.split(",")]:
        for bps in [int(x) for x in a.bps.split(",")]:
            f.set_search_variant(lanes, 0, bps)
            for it in range(3):
                f.profile_reset()
                ix.count_device(d_q.data_ptr(), d_off.data_ptr(), a.nq, d_cnt.data_ptr(), st)
                torch.cuda.synchronize()
                p = f.profile_get()
            steps = a.nq * (a.qlen - a.k)
            print(f"lanes={lanes} bps={bps}: search {p['search_ms']:.2f} ms pack {p['pack_ms']:.2f} ms  "
                  f"{a.nq/p['search_ms']/1e3:.1f} M reads/s  {steps/p['search_ms']/1e6:.2f} G LF-steps/s  "
                  f"alg {steps*(104 if a.alphabet == 0 else 168)/p['search_ms']/1e6:.0f} GB/s", flush=True)
            results.append({"lanes": lanes, "bps": bps, **p})
    f.set_search_variant(0)
    ix.device_check(st)
    cnt = d_cnt.cpu().numpy()
    print("count histogram:", np.unique(cnt, return_counts=True)[0][:5], flush=True)
    if a.check:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from oracle import pyoracle as po
        orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len,
                                        parts.blocks, parts.prefix_sums, parts.sa_words)
        qb = d_q[: a.check * a.qlen].cpu().numpy()
        qo = np.arange(a.check + 1, dtype=np.uint64) * np.uint64(a.qlen)
        t2 = time.time()
        want, st_or = orc.count_batch(qb, qo)
        dt = time.time() - t2
        print(f"oracle: {a.check} reads in {dt:.2f}s on {os.cpu_count()} threads = {a.check/dt/1e3:.1f} k reads/s; "
              f"parity {'OK' if np.array_equal(want, cnt[:a.check].astype(np.uint64)) else 'MISMATCH'}; stats {st_or}", flush=True)
        # locate
        nq2, ql2 = a.locate_nq, a.locate_qlen
        d_q2 = torch.empty(nq2 * ql2, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(a.alphabet, a.n, 3, nq2, ql2, 5, d_q2.data_ptr())
        d_off2 = torch.arange(0, nq2 + 1, dtype=torch.int64, device="cuda") * ql2
        d_hoff = torch.zeros(nq2 + 1, dtype=torch.int64, device="cuda")
        for variant in (1, 0):      # LF-walk, then the unsampled-SA gather (left selected for the parity check)
            f.set_locate_variant(variant)
            for it in range(3):
                f.profile_reset()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ptr, nh = ix.locate_device(d_q2.data_ptr(), d_off2.data_ptr(), nq2, d_hoff.data_ptr(), stream=st)
                e1.record()
                torch.cuda.synchronize()
                p = f.profile_get()
                if it < 2 or variant == 1:
                    ix.device_free(ptr)
            print(f"locate variant {variant} ({'LF-walk' if variant else 'unsampled-SA gather'}): {nh} hits; total "
                  f"{e0.elapsed_time(e1):.2f} ms; search {p['search_ms']:.2f} ms pass2 {p['walk_ms']:.3f} ms "
                  f"{nh/p['walk_ms']/1e3:.1f} M hits/s in pass 2", flush=True)
        ncheck = min(a.check, nq2)
        qb2 = d_q2[: ncheck * ql2].cpu().numpy()
        qo2 = np.arange(ncheck + 1, dtype=np.uint64) * np.uint64(ql2)
        woff, whits, st2 = orc.locate_batch(qb2, qo2)
        hoff = d_hoff.cpu().numpy().astype(np.uint64)
        nh_check = int(hoff[ncheck])
        import ctypes as C
        buf = torch.empty(nh_check * 2, dtype=torch.int64, device="cuda")
        C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), C.c_size_t(nh_check * 16), 3)
        got = buf.cpu().numpy().astype(np.uint64).reshape(-1, 2)
        ok = np.array_equal(hoff[: ncheck + 1], woff) and np.array_equal(got, whits)
        print(f"locate parity on {ncheck} queries: {'OK' if ok else 'MISMATCH'}; oracle stats {st2}", flush=True)
        ix.device_free(ptr)
    print(f"total {time.time()-t0:.1f}s")


if __name__ == "__main__":
    main()
