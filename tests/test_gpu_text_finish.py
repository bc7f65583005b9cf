"""GPU parity of the nucleotide count accelerator "finish in the text" (awry_set_count_variant, layout.cuh:
IndexView::rtext): once an interval is one row wide the count kernel reads SA[row] and compares the rest of the
query with the text instead of stepping on.  Counts must equal backward search to the last symbol (variant 1)
and the oracle on every kind of query that can reach the comparison: matches, mismatches at every distance from
the one-row point, queries that would start before the text, queries across record delimiters (literal N,
fm_index.rs:148-153), N inside the text, N inside the query (the scalar kernel), lengths 1 .. 600 (the
shared-memory ring holds 256 symbols)."""
import numpy as np
import pytest

from conftest import device_from_parts, oracle_from_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def multi(fx):
    r = np.random.default_rng(11)
    recs = []
    for i, n in enumerate([70_000, 1, 130_000, 40, 99_999]):
        t = bytearray(bytes(fx.gen_text(0, n, 20 + i)))
        if n > 1000:                                   # N runs and isolated N inside the records
            for _ in range(6):
                t[int(r.integers(0, n))] = ord("N")
            p = int(r.integers(0, n - 300))
            t[p:p + 120] = b"N" * 120
        recs.append(bytes(t))
    text, starts = fx.concat_records(recs, 0)
    return fx.build_parts(text, 0, ratio=8, kmer_len=9, seq_starts=starts)


@pytest.fixture(scope="module")
def multi_dev(multi):
    ix = device_from_parts(multi)
    yield ix
    ix.close()


def queries_of_every_kind(parts, n, seed, max_len=600):
    from awry_b200 import fm_index as f
    r = np.random.default_rng(seed)
    t = bytes(parts.text)
    starts = [int(s) for s in parts.seq_starts]
    qs = []
    for i in range(n):
        kind = i % 16
        ln = int(r.integers(1, max_len)) if kind != 15 else int(r.integers(1, 40))
        if kind == 0:                                   # hangs over the START of the text: a random head + text[0:m]
            m = int(r.integers(1, min(ln, 200) + 1))
            head = bytes(r.choice(np.frombuffer(b"ACGT", np.uint8), ln))
            q = bytearray(head + t[:m])
        elif kind == 1:                                 # the very end of the text
            q = bytearray(t[len(t) - min(ln, len(t)):])
        elif kind == 2 and len(starts) > 1:             # across a record delimiter (contains the literal N)
            s = starts[int(r.integers(1, len(starts)))]
            a = max(0, s - int(r.integers(1, 100)))
            q = bytearray(t[a:a + ln])
        elif kind == 3 and len(starts) > 1:             # across a delimiter with the N replaced: never a substring there
            s = starts[int(r.integers(1, len(starts)))]
            a = max(0, s - int(r.integers(1, 100)))
            q = bytearray(t[a:a + ln].replace(b"N", b"A"))
        else:
            p = int(r.integers(0, len(t) - ln))
            q = bytearray(t[p:p + ln])
        if kind in (4, 5, 6) and len(q) > 0:            # one substitution, anywhere (incl. the first / last symbol)
            at = [int(r.integers(0, len(q))), 0, len(q) - 1][kind - 4]
            q[at] = ord("ACGT"[(b"ACGT".find(bytes([q[at]])) + 1) % 4]) if q[at] in b"ACGT" else ord("A")
        if kind == 7:
            q = bytearray(bytes(q).lower())
        if kind == 8 and len(q) > 0:
            q[int(r.integers(0, len(q)))] = ord("NRnyk"[int(r.integers(0, 5))])
        qs.append(bytes(q))
    return f.pack_queries(qs)


def test_counts_equal_backward_search_and_the_oracle(po, multi, multi_dev):
    from awry_b200 import fm_index as f
    assert multi_dev.device_bytes()["text"] > 0 and multi_dev.device_bytes()["full_sa"] > 0
    orc = oracle_from_parts(po, multi)
    qb, qo = queries_of_every_kind(multi, 40_000, seed=3)
    want, _ = orc.count_batch(qb, qo)
    assert (want == 1).sum() > 5_000 and (want == 0).sum() > 5_000 and (want > 1).sum() > 100
    try:
        f.set_count_variant(1)
        plain = multi_dev.count_packed(qb, qo)
    finally:
        f.set_count_variant(0)
    assert np.array_equal(plain, want)
    try:
        for kernel in (0, 80, 81, 82, 83, 84):              # wave kernel, refilling kernel, state-machine kernel (1 / 2 slots)
            f.set_search_variant(kernel)
            assert np.array_equal(multi_dev.count_packed(qb, qo), want), kernel
    finally:
        f.set_search_variant(0)
    crumbs, exc = f.host_pack_dna(qb)                       # the pre-packed entry point runs the same kernel
    assert np.array_equal(multi_dev.count_prepacked(crumbs, qo, exc), want)


def test_locate_through_the_text_equals_the_oracle(po, multi, multi_dev):
    """locate pass 1 stores a query finished in the text as its text position (CNT_AT_TEXT_POS) and pass 2 -- the
    gather from the unsampled array -- passes it through: same CSR offsets and hits as the oracle, through the
    owned-buffer call, the caller-buffer call (hits written by the gather kernel into pinned memory), the
    pre-packed call, sorted output, and with the walks (locate variants 1 / 2: pass 1 then keeps rows)"""
    import torch
    from awry_b200 import fm_index as f
    orc = oracle_from_parts(po, multi)
    qb, qo = queries_of_every_kind(multi, 30_000, seed=5)
    woff, whits, _ = orc.locate_batch(qb, qo)
    soff, shits, _ = orc.locate_batch(qb, qo, sorted_hits=True)
    assert (np.diff(woff.astype(np.int64)) == 1).sum() > 4_000

    def pinned(shape):
        return torch.zeros(shape, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)

    for count_variant in (0, 1):
        f.set_count_variant(count_variant)
        try:
            for lv in (0, 1, 2):
                f.set_locate_variant(lv)
                off, hits = multi_dev.locate_packed(qb, qo)
                assert np.array_equal(off, woff) and np.array_equal(hits, whits), (count_variant, lv)
            f.set_locate_variant(0)
            off, hits = multi_dev.locate_packed(qb, qo, sorted_hits=True)
            assert np.array_equal(off, soff) and np.array_equal(hits, shits)
            hoff, hh = pinned(len(qo)), pinned((len(whits) + 3, 2))
            pq, pqo = torch.from_numpy(qb).pin_memory().numpy(), pinned(len(qo))
            pqo[:] = qo
            n = multi_dev.locate_packed_into(pq, pqo, hoff, hh)
            assert n == len(whits) and np.array_equal(hoff, woff) and np.array_equal(hh[:n], whits)
            crumbs, exc = f.host_pack_dna(qb)
            hoff[:] = 0
            hh[:] = 0
            n = multi_dev.locate_prepacked_into(crumbs, qo, exc, hoff, hh)
            assert n == len(whits) and np.array_equal(hoff, woff) and np.array_equal(hh[:n], whits)
        finally:
            f.set_count_variant(0)
            f.set_locate_variant(0)


def test_every_length_and_every_mismatch_position(po, fx):
    """one read of every length 1 .. 330 from one place, exact and with the substitution walked over every
    position of a 300-symbol read: the comparison starts at every alignment of query words against text words"""
    from awry_b200 import fm_index as f
    text = fx.gen_text(0, 200_003, 77)
    parts = fx.build_parts(text, 0, ratio=4, kmer_len=6)
    orc = oracle_from_parts(po, parts)
    t = bytes(parts.text)
    qs = []
    for base in (0, 1, 7, 64, 65, 100_001, len(t) - 330):
        for ln in range(1, 331):
            qs.append(t[base:base + ln])
            qs.append(t[len(t) - ln:])
        read = bytearray(t[base:base + 300])
        for at in range(300):
            q = bytearray(read)
            q[at] = ord("ACGT"[(b"ACGT".find(bytes([q[at]])) + 2) % 4])
            qs.append(bytes(q))
    qb, qo = f.pack_queries(qs)
    want, _ = orc.count_batch(qb, qo)
    ix = device_from_parts(parts)
    try:
        assert ix.device_bytes()["text"] > 0
        assert np.array_equal(ix.count_packed(qb, qo), want)
        f.set_count_variant(1)
        assert np.array_equal(ix.count_packed(qb, qo), want)
    finally:
        f.set_count_variant(0)
        ix.close()


def test_long_reads_are_compared_out_of_global_memory(po, multi, multi_dev):
    """reads longer than the 256-symbol ring (1-8 kbp): the wave kernel hands them to the refilling kernel, which
    compares against the packed words in global memory; exact, with a substitution near either end, and hanging
    over the start of the text"""
    from awry_b200 import fm_index as f
    orc = oracle_from_parts(po, multi)
    r = np.random.default_rng(21)
    t = bytes(multi.text)
    qs = []
    for i in range(1500):
        ln = int(r.integers(257, 8000))
        p = int(r.integers(0, len(t) - ln)) if i % 7 else 0
        q = bytearray(t[p:p + ln])
        if i % 3 == 1:
            at = [int(r.integers(0, ln)), 0, ln - 1, int(r.integers(0, 20)), ln - 1 - int(r.integers(0, 300))][i % 5]
            q[at] = ord("ACGT"[(b"ACGT".find(bytes([q[at]])) + 1) % 4]) if q[at] in b"ACGT" else ord("C")
        if i % 11 == 0:
            q = bytearray(bytes(r.choice(np.frombuffer(b"ACGT", np.uint8), 5)) + t[:ln])   # starts before the text
        qs.append(bytes(q))
    qb, qo = f.pack_queries(qs)
    want, _ = orc.count_batch(qb, qo)
    assert (want >= 1).sum() > 300 and (want == 0).sum() > 300
    for kernel in (0, 80):
        f.set_search_variant(kernel)
        try:
            assert np.array_equal(multi_dev.count_packed(qb, qo), want), kernel
        finally:
            f.set_search_variant(0)
    woff, whits, _ = orc.locate_batch(qb, qo)
    off, hits = multi_dev.locate_packed(qb, qo)
    assert np.array_equal(off, woff) and np.array_equal(hits, whits)


def test_protein_counts_and_hits_through_the_text(po, fx):
    """the protein kernel finishes one-row intervals the same way (one byte per symbol): peptides of every
    length 1 .. 200 (the ring holds 128 symbols), a substitution at every position of a 40-residue peptide,
    peptides hanging over the start of the text and across record delimiters (X), X in the query; counts and
    locate lists equal the oracle with the step on and off"""
    from awry_b200 import fm_index as f
    recs = [bytes(fx.gen_text(1, n, 40 + i)) for i, n in enumerate([60_000, 3, 90_000, 25])]
    text, starts = fx.concat_records(recs, 1)
    parts = fx.build_parts(text, 1, ratio=8, kmer_len=3, seq_starts=starts)
    orc = oracle_from_parts(po, parts)
    t = bytes(parts.text)
    r = np.random.default_rng(8)
    letters = b"ACDEFGHIKLMNPQRSTVWY"
    qs = []
    for base in (0, 1, 5, 31, 32, 33, 70_001, len(t) - 200):
        for ln in range(1, 201):
            qs.append(t[base:base + ln])
            qs.append(t[len(t) - ln:])
        pep = bytearray(t[base:base + 40])
        for at in range(40):
            q = bytearray(pep)
            q[at] = letters[(letters.find(bytes([q[at]])) + 3) % 20] if q[at] in letters else ord("A")
            qs.append(bytes(q))
    for i in range(4000):
        ln = int(r.integers(1, 60))
        kind = i % 8
        if kind == 0:
            q = bytes(r.choice(np.frombuffer(letters, np.uint8), int(r.integers(1, 6)))) + t[:ln]       # before the text
        elif kind == 1:
            s0 = int(starts[int(r.integers(1, len(starts)))])
            a = max(0, s0 - int(r.integers(1, 20)))
            q = t[a:a + ln]                                                                       # across a delimiter (X)
        else:
            p0 = int(r.integers(0, len(t) - ln))
            q = bytearray(t[p0:p0 + ln])
            if kind == 2:
                q[int(r.integers(0, ln))] = ord("X")
            if kind == 3:
                at = int(r.integers(0, ln))
                q[at] = letters[(letters.find(bytes([q[at]])) + 1) % 20] if q[at] in letters else ord("C")
            q = bytes(q)
        qs.append(q)
    qb, qo = f.pack_queries(qs)
    want, _ = orc.count_batch(qb, qo)
    woff, whits, _ = orc.locate_batch(qb, qo)
    assert (want == 1).sum() > 2_000 and (want == 0).sum() > 500
    ix = device_from_parts(parts)
    try:
        assert ix.device_bytes()["text"] > 0
        for count_variant in (0, 1):
            f.set_count_variant(count_variant)
            for kernel in (0, 80):                          # wave kernel / refilling kernel
                f.set_search_variant(kernel)
                assert np.array_equal(ix.count_packed(qb, qo), want), (count_variant, kernel)
                off, hits = ix.locate_packed(qb, qo)
                assert np.array_equal(off, woff) and np.array_equal(hits, whits), (count_variant, kernel)
    finally:
        f.set_count_variant(0)
        f.set_search_variant(0)
        ix.close()


def test_without_the_text_the_kernel_steps_to_the_end(po, fx, monkeypatch):
    from awry_b200 import fm_index as f
    monkeypatch.setenv("AWRY_B200_TEXT", "0")
    text = fx.gen_text(0, 50_000, 5)
    parts = fx.build_parts(text, 0, ratio=8, kmer_len=8)
    ix = device_from_parts(parts)
    try:
        assert ix.device_bytes()["text"] == 0
        qb, qo, _ = fx.gen_substring_queries(text, 2_000, 100, 9)
        want, _ = oracle_from_parts(po, parts).count_batch(qb, qo)
        assert np.array_equal(ix.count_packed(qb, qo), want)
    finally:
        ix.close()
