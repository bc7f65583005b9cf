"""One process, N GPUs: the in-library multi-replica path (what the Rust facade would use).
Builds the cfg2 index, replicates it over all visible GPUs (NVLink fan-out) and times
awry_count_batch on pinned host reads split across replicas by the library."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from awry_b200 import FmIndex  # noqa: E402
from fixtures import pyfixture_gpu as fxg  # noqa: E402


def main():
    n_dev = torch.cuda.device_count()
    n, nq_per, L = 3_100_000_000, 10_000_000, 150
    parts, phases = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
    for devs in ([0], list(range(min(2, n_dev))), list(range(min(4, n_dev))), list(range(n_dev))):
        if len(devs) > n_dev or (len(devs) > 1 and devs == [0]):
            continue
        t0 = time.time()
        ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                parts.prefix_sums, parts.sa_words, devices=devs)
        t_load = time.time() - t0
        nq = nq_per * len(devs)
        torch.cuda.set_device(0)
        d = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, n, 3, nq, L, 4, d.data_ptr())
        h_q = torch.empty(nq * L, dtype=torch.uint8, pin_memory=True)
        h_q.copy_(d)
        del d
        h_off = (torch.arange(0, nq + 1, dtype=torch.int64) * L).pin_memory()
        h_cnt = torch.zeros(nq, dtype=torch.int64, pin_memory=True)
        qb, qo, out = h_q.numpy(), h_off.numpy().view(np.uint64), h_cnt.numpy().view(np.uint64)
        for _ in range(2):
            ix.count_packed(qb, qo, out=out)
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            ix.count_packed(qb, qo, out=out)
        dt = (time.perf_counter() - t0) / reps
        print(f"devices={len(devs)}: replicate {t_load:.2f}s; {nq} reads in {dt*1e3:.1f} ms = {nq/dt/1e6:.1f} M reads/s "
              f"(min count {int(out.min())})", flush=True)
        ix.close()
        del h_q, h_off, h_cnt


if __name__ == "__main__":
    main()
