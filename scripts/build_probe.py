"""GPU index construction at BASELINE scale through the product entry points: a 3.1 Gbp single-record FASTA
(80 columns) -> awry_build_index_file (`.awry` v1 with the reference-style k = 13 table section) and
awry_index_build (straight to a searchable index); the result is checked against the index built from the
same text through awry_build_parts + from_parts (which tests/test_gpu_build.py pins to the CPU builder)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from awry_b200 import FmBuildArgs, FmIndex, fm_index as f  # noqa: E402
from fixtures import pyfixture as fx, pyfixture_gpu as fxg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=3_100_000_000)
    ap.add_argument("--k", type=int, default=13)
    ap.add_argument("--ratio", type=int, default=8)
    ap.add_argument("--dir", default="/dev/shm")
    a = ap.parse_args()
    t0 = time.time()
    d_text = torch.empty(a.n, dtype=torch.uint8, device="cuda")
    fxg.lib().fxg_gen_text_device(0, a.n, 3, d_text.data_ptr(), 0)
    text = d_text.cpu().numpy()
    del d_text
    fa = os.path.join(a.dir, "awry_probe.fa")
    rows, rem = divmod(a.n, 80)
    with open(fa, "wb") as fh:
        fh.write(b">synthetic\n")
        body = np.empty((rows, 81), dtype=np.uint8)
        body[:, :80] = text[: rows * 80].reshape(rows, 80)
        body[:, 80] = 10
        body.tofile(fh)
        if rem:
            fh.write(text[rows * 80:].tobytes() + b"\n")
    del body
    print(f"FASTA {os.path.getsize(fa)/1e9:.2f} GB written in {time.time()-t0:.1f}s", flush=True)
    out = os.path.join(a.dir, "awry_probe.awry")
    try:
        t1 = time.time()
        f.build_index_file(fa, out, 0, suffix_array_compression_ratio=a.ratio, lookup_table_kmer_len=a.k)
        t2 = time.time()
        print(f"awry_build_index_file: {t2-t1:.2f}s -> {os.path.getsize(out)/1e9:.2f} GB .awry file", flush=True)
        ix = FmIndex.new(FmBuildArgs(fa, suffix_array_compression_ratio=a.ratio, lookup_table_kmer_len=a.k, alphabet=0))
        t3 = time.time()
        print(f"awry_index_build (FASTA -> searchable index incl. seed table, pair index, unsampled SA): {t3-t2:.2f}s", flush=True)
        blocks, prefix, sa_words, phases = f.build_parts(0, text, sa_ratio=a.ratio)
        print("awry_build_parts phases (s):", {k: round(v, 2) for k, v in phases.items()}, flush=True)
        ix2 = FmIndex.load(out)
        t4 = time.time()
        nq, L = 2_000_000, 50
        d_q = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, a.n, 3, nq, L, 11, d_q.data_ptr())
        qb = d_q.cpu().numpy()
        qo = np.arange(nq + 1, dtype=np.uint64) * np.uint64(L)
        c1, c2 = ix.count_packed(qb, qo), ix2.count_packed(qb, qo)
        o1, h1 = ix.locate_packed(qb[: 100_000 * L], qo[:100_001])
        o2, h2 = ix2.locate_packed(qb[: 100_000 * L], qo[:100_001])
        ok = np.array_equal(c1, c2) and int(c1.min()) >= 1 and np.array_equal(h1, h2) and np.array_equal(o1, o2)
        win = fx.gen_text_windows(0, 3, h1[:, 1], L)
        ok = ok and np.array_equal(win, qb[: 100_000 * L].reshape(-1, L)[np.searchsorted(o1, np.arange(len(h1)), side="right") - 1])
        # the file's arrays == the arrays of awry_build_parts
        with open(out, "rb") as fh:
            fh.seek(43)
            fb = np.fromfile(fh, dtype=np.uint64, count=len(blocks))
            fp = np.fromfile(fh, dtype=np.uint64, count=7)
            fs = np.fromfile(fh, dtype=np.uint64, count=len(sa_words))
        ok = ok and np.array_equal(fb, blocks) and np.array_equal(fp, prefix) and np.array_equal(fs, sa_words)
        print(f"load of the built file {t4-t3:.2f}s; built index vs loaded file vs text: {'OK' if ok else 'MISMATCH'}", flush=True)
    finally:
        for p in (fa, out):
            if os.path.exists(p):
                os.remove(p)


if __name__ == "__main__":
    main()
