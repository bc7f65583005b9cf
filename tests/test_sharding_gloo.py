"""CPU tests of the multi-rank path: world size 2 over gloo.  The per-rank "engine" is the oracle
(no GPU here); what is under test is the partitioning, order preservation and the gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from awry_b200 import sharding
    from fixtures import pyfixture as fx
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    text = fx.gen_text(0, 30_000, 21)
    parts = fx.build_parts(text, 0, ratio=4, kmer_len=5)
    orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                    parts.prefix_sums, parts.sa_words)
    rng = np.random.default_rng(7)
    t = bytes(text)
    qs = []
    for _ in range(999):                       # ragged lengths, some absent
        n = int(rng.integers(1, 90))
        p = int(rng.integers(0, len(t) - n))
        q = bytearray(t[p:p + n])
        if rng.random() < 0.2:
            q[0] = ord("N")
        qs.append(bytes(q))
    qb, qo = po.pack_queries(qs)
    counts = sharding.sharded_count(qb, qo, lambda b, o: orc.count_batch(b, o)[0])
    loc = sharding.sharded_locate(qb, qo, lambda b, o: orc.locate_batch(b, o)[:2])
    if rank == 0:
        want, _ = orc.count_batch(qb, qo)
        woff, whits, _ = orc.locate_batch(qb, qo)
        ok = (np.array_equal(counts, want) and np.array_equal(loc[0], woff) and np.array_equal(loc[1], whits))
        open(out_path, "w").write("ok" if ok else "mismatch")
    else:
        assert counts is None and loc is None
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_count_and_locate_gloo(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_split_by_bytes_properties():
    from awry_b200.sharding import rank_slice, split_by_bytes
    rng = np.random.default_rng(1)
    for nq in (0, 1, 3, 4, 17, 1000):
        lens = rng.integers(1, 300, nq)
        qoff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        for parts in (1, 2, 3, 8):
            r = split_by_bytes(qoff, parts)
            assert len(r) == parts and r[0][0] == 0 and r[-1][1] == nq
            assert all(a[1] == b[0] for a, b in zip(r, r[1:])) and all(lo <= hi for lo, hi in r)
            if nq >= 2 * parts and parts > 1:
                sizes = [int(qoff[h] - qoff[l]) for l, h in r]
                assert max(sizes) - min(sizes) <= 2 * 300
    qoff = np.array([0, 5, 9, 20, 21, 40, 44, 60, 61], dtype=np.uint64)
    qb = np.arange(61, dtype=np.uint8)
    sb, so, (lo, hi) = rank_slice(qb, qoff, 1, 2)
    assert so[0] == 0 and len(so) == hi - lo + 1 and len(sb) == int(so[-1])
    assert np.array_equal(sb, qb[int(qoff[lo]):int(qoff[hi])])
