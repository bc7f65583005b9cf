"""A/B of search-kernel variants on the cfg2 workload: interleaved repeats, min and median per variant."""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from awry_b200 import FmIndex, fm_index as f  # noqa: E402
from fixtures import pyfixture_gpu as fxg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="8:0,9:0,8:4,9:4")
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--burst", type=int, default=1, help="launches back to back per measurement (sustained clocks)")
    ap.add_argument("--n", type=int, default=3_100_000_000)
    ap.add_argument("--nq", type=int, default=10_000_000)
    ap.add_argument("--qlen", type=int, default=150)
    ap.add_argument("--mut-ppm", type=int, default=0, help="reads per million carrying one substitution (early exit)")
    a = ap.parse_args()
    parts, _ = fxg.build_parts(0, a.n, 3, ratio=8, kmer_len=13)
    os.environ["AWRY_B200_FULL_SA"] = "0"
    ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words)
    d_q = torch.empty(a.nq * a.qlen, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(0, a.n, 3, a.nq, a.qlen, 4, d_q.data_ptr(), mut_ppm=a.mut_ppm)
    d_off = torch.arange(0, a.nq + 1, dtype=torch.int64, device="cuda") * a.qlen
    d_cnt = torch.zeros(a.nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    f.profile_enable(True)
    variants = [tuple(int(x) for x in v.split(":")) for v in a.variants.split(",")]
    times = {v: [] for v in variants}
    ref = None
    for rep in range(a.reps):
        for v in variants:
            f.set_search_variant(v[0], 0, v[1])
            f.profile_reset()
            for _ in range(a.burst):
                ix.count_device(d_q.data_ptr(), d_off.data_ptr(), a.nq, d_cnt.data_ptr(), st)
            torch.cuda.synchronize()
            times[v].append(f.profile_get()["search_ms"] / a.burst)
            ref = d_cnt.clone() if ref is None else ref
            assert torch.equal(d_cnt, ref), "variants disagree"
    for v in variants:
        t = times[v][1:]
        print(f"lanes={v[0]} bps={v[1]}: min {min(t):.2f} ms  median {statistics.median(t):.2f} ms  all {[round(x, 2) for x in times[v]]}", flush=True)


if __name__ == "__main__":
    main()
