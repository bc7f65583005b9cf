"""Randomised differential test: small random indexes (both alphabets, 1-5 records, text lengths from 1 symbol
up, SA ratios 1-40, seed-table lengths 1-9, ambiguity symbols) and ragged random queries, CUDA path vs the CPU
oracle: counts, ranges, locate lists in BWT order and sorted, both locate variants, host packing on and off."""
import numpy as np
import pytest

from conftest import device_from_parts, oracle_from_parts

pytestmark = pytest.mark.gpu


def _case(fx, rng, alphabet):
    letters = b"ACGT" if alphabet == 0 else b"ACDEFGHIKLMNPQRSTVWY"
    amb = b"NRY" if alphabet == 0 else b"XBZ"
    n_rec = int(rng.integers(1, 6))
    lens = [int(rng.choice([1, 2, 3, 7, 31, 127, 128, 129, 255, 256, 257, 1000, 4000])) for _ in range(n_rec)]
    skew = rng.random() < 0.3          # low-entropy text: wide intervals, straddling blocks
    recs = []
    for ln in lens:
        p = np.array([0.7, 0.1, 0.1, 0.1]) if (skew and alphabet == 0) else None
        idx = rng.choice(len(letters), ln, p=p) if p is not None else rng.integers(0, len(letters), ln)
        r = bytearray(np.frombuffer(letters, dtype=np.uint8)[idx].tobytes())
        for _ in range(int(rng.integers(0, 3))):
            r[int(rng.integers(0, ln))] = amb[int(rng.integers(0, len(amb)))]
        recs.append(bytes(r))
    text, starts = fx.concat_records(recs, alphabet)
    ratio = int(rng.choice([1, 2, 3, 4, 7, 8, 16, 32, 40]))
    k = int(rng.integers(1, 10 if alphabet == 0 else 5))
    parts = fx.build_parts(text, alphabet, ratio=ratio, kmer_len=k, seq_starts=starts)
    t = bytes(text)
    qs = []
    for _ in range(300):
        ln = int(rng.choice([1, 2, 3, 5, 8, 12, 13, 14, 31, 32, 33, 64, 150, 300]))
        if rng.random() < 0.7 and len(t) >= 1:
            p0 = int(rng.integers(0, len(t)))
            q = bytearray(t[p0:p0 + ln])
        else:
            q = bytearray(np.frombuffer(letters, dtype=np.uint8)[rng.integers(0, len(letters), ln)].tobytes())
        if not q:
            q = bytearray(letters[:1])
        if rng.random() < 0.1:
            q[int(rng.integers(0, len(q)))] = amb[0]
        if rng.random() < 0.1:
            q = bytearray(bytes(q).lower())
        qs.append(bytes(q))
    return parts, qs


@pytest.mark.parametrize("seed", range(40))
def test_random_index_and_queries(fx, po, seed):
    from awry_b200 import fm_index as f
    rng = np.random.default_rng(1000 + seed)
    alphabet = seed % 2
    parts, qs = _case(fx, rng, alphabet)
    orc = oracle_from_parts(po, parts)
    # pad the batch so that the host-packed path (>= 4096 query bytes) is exercised as well
    qb, qo = f.pack_queries(qs * 3)
    want, _ = orc.count_batch(qb, qo)
    woff, whits, _ = orc.locate_batch(qb, qo)
    soff, shits, _ = orc.locate_batch(qb, qo, sorted_hits=True)
    with device_from_parts(parts) as ix:
        try:
            for pack in (1, 0):
                f.set_host_pack(pack)
                assert np.array_equal(ix.count_packed(qb, qo), want), (seed, pack)
                rng_ = ix.search_packed(qb, qo)
                cnt = np.where(rng_[:, 0] > rng_[:, 1], 0, rng_[:, 1] - rng_[:, 0] + 1)
                assert np.array_equal(cnt.astype(np.uint64), want), (seed, pack)
                for variant in (0, 1):
                    f.set_locate_variant(variant)
                    off, hits = ix.locate_packed(qb, qo)
                    assert np.array_equal(off, woff) and np.array_equal(hits, whits), (seed, pack, variant)
                    off2, hits2 = ix.locate_packed(qb, qo, sorted_hits=True)
                    assert np.array_equal(off2, soff) and np.array_equal(hits2, shits), (seed, pack, variant)
        finally:
            f.set_host_pack(-1)
            f.set_locate_variant(0)
