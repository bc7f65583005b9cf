"""Streaming reads-file front-end at scale: a synthetic FASTQ of N x L-bp reads (exact substrings of the
3.1 Gbp synthetic text) counted straight from the file (awry_count_reads_file: pread ring -> H2D -> device
parser -> pack -> search), against the same reads through the device-resident entry point."""
import argparse
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
    ap.add_argument("--nq", type=int, default=10_000_000)
    ap.add_argument("--qlen", type=int, default=150)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--chunks", default="64,256")
    ap.add_argument("--devices", type=int, default=1, help="replicas behind the handle (the file is cut into as many segments)")
    ap.add_argument("--copies", type=int, default=1, help="write the reads this many times into the file")
    a = ap.parse_args()
    parts, _ = fxg.build_parts(0, a.n, 3, ratio=8, kmer_len=13)
    os.environ.setdefault("AWRY_B200_FULL_SA", "1")
    os.environ.setdefault("AWRY_B200_LEAN_SA", "0")
    ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words, devices=list(range(a.devices)))
    d_q = torch.empty(a.nq * a.qlen, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(0, a.n, 3, a.nq, a.qlen, 4, d_q.data_ptr())
    d_off = torch.arange(0, a.nq + 1, dtype=torch.int64, device="cuda") * a.qlen
    d_cnt = torch.zeros(a.nq, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ix.count_device(d_q.data_ptr(), d_off.data_ptr(), a.nq, d_cnt.data_ptr(), st)
    ix.device_check(st)
    want = d_cnt.cpu().numpy().view(np.uint64)
    # FASTQ: "@r\n" + seq + "\n+\n" + qual + "\n"
    t0 = time.time()
    rec = np.empty((a.nq, 3 + a.qlen + 3 + a.qlen + 1), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
    rec[:, 3:3 + a.qlen] = d_q.cpu().numpy().reshape(a.nq, a.qlen)
    rec[:, 3 + a.qlen:6 + a.qlen] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 6 + a.qlen:6 + 2 * a.qlen] = ord("I")
    rec[:, -1] = ord("\n")
    path = os.path.join(a.dir, "awry_probe_reads.fq")
    with open(path, "wb") as fh:
        for _ in range(a.copies):
            rec.tofile(fh)
    size = os.path.getsize(path)
    del rec
    want = np.tile(want, a.copies)
    a.nq *= a.copies
    print(f"wrote {path}: {size/1e9:.2f} GB, {a.nq} reads in {time.time()-t0:.1f}s", flush=True)
    try:
        for chunk_mb in [int(x) for x in a.chunks.split(",")]:
            os.environ["AWRY_B200_READS_CHUNK"] = str(chunk_mb << 20)
            for it in range(3):
                t1 = time.perf_counter()
                got = ix.count_reads_file(path)
                dt = time.perf_counter() - t1
            ok = np.array_equal(got, want)
            print(f"count_reads_file chunk {chunk_mb} MiB: {dt*1e3:.0f} ms = {a.nq/dt/1e6:.1f} M reads/s, "
                  f"{size/dt/1e9:.2f} GB/s of FASTQ; parity vs device-resident counts {'OK' if ok else 'MISMATCH'}", flush=True)
        if a.devices > 1:
            os.environ["AWRY_B200_READS_REPLICAS"] = "1"
            t1 = time.perf_counter()
            got = ix.count_reads_file(path)
            dt = time.perf_counter() - t1
            print(f"same handle, replica 0 only: {dt*1e3:.0f} ms = {a.nq/dt/1e6:.1f} M reads/s; parity {'OK' if np.array_equal(got, want) else 'MISMATCH'}", flush=True)
            del os.environ["AWRY_B200_READS_REPLICAS"]
        for it in range(3):          # (the first call loads the locate kernels and grows the buffers)
            t1 = time.perf_counter()
            off, hits = ix.locate_reads_file(path)
            dt = time.perf_counter() - t1
        print(f"locate_reads_file: {len(hits)} hits in {dt*1e3:.0f} ms = {a.nq/dt/1e6:.1f} M reads/s", flush=True)
    finally:
        os.remove(path)


if __name__ == "__main__":
    main()
