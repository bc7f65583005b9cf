"""GPU parity tests of the round-2 entry points: pre-packed (2-bit) reads, the sync-free locate pass 2,
offset validation, and one batch split over several replicas inside one process."""
import numpy as np
import pytest

from conftest import device_from_parts, mixed_queries, oracle_from_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dna(fx):
    text = fx.gen_text(0, 300_000, 5)
    return fx.build_parts(text, 0, ratio=8, kmer_len=9)


@pytest.fixture(scope="module")
def dna_dev(dna):
    ix = device_from_parts(dna)
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def dna_or(po, dna):
    return oracle_from_parts(po, dna)


def pinned(shape, dtype=np.uint64):
    import torch
    t = torch.zeros(shape, dtype=torch.int64 if dtype == np.uint64 else torch.uint8, pin_memory=True)
    return t.numpy().view(dtype)


def ragged_queries(dna, n, seed, with_n=True):
    from awry_b200 import fm_index as f
    r = np.random.default_rng(seed)
    t = bytes(dna.text)
    qs = []
    for i in range(n):
        ln = int(r.integers(1, 180))
        p = int(r.integers(0, len(t) - ln))
        q = bytearray(t[p:p + ln])
        if with_n and i % 53 == 0:
            q[int(r.integers(0, ln))] = ord("NRnyk"[int(r.integers(0, 5))])
        if i % 71 == 0:
            q = bytearray(bytes(q).lower())
        if i % 5 == 0:
            q[int(r.integers(0, ln))] = ord("ACGT"[int(r.integers(0, 4))])       # early exits
        qs.append(bytes(q))
    return f.pack_queries(qs)


@pytest.mark.parametrize("pin", [False, True], ids=["pageable", "pinned"])
def test_prepacked_reads_count_and_locate(dna, dna_dev, dna_or, pin):
    """awry_count_batch_packed2 / awry_locate_batch_packed2 == the ASCII entry points == the oracle"""
    from awry_b200 import fm_index as f
    qb, qo = ragged_queries(dna, 20_000, seed=1)
    crumbs, exc = f.host_pack_dna(qb)
    assert len(exc) > 100
    want, _ = dna_or.count_batch(qb, qo)
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    if pin:
        pc, po_, pe = pinned(len(crumbs) + 8, np.uint8), pinned(len(qo)), pinned(len(exc))
        pc[:len(crumbs)] = crumbs
        po_[:] = qo
        pe[:] = exc
        crumbs, qo2, exc = pc, po_, pe
    else:
        qo2 = qo
    got = dna_dev.count_prepacked(crumbs, qo2, exc)
    assert np.array_equal(got, want)
    assert np.array_equal(dna_dev.count_packed(qb, qo), want)
    hoff, hits = pinned(len(qo)), pinned((len(whits) + 5, 2))
    n = dna_dev.locate_prepacked_into(crumbs, qo2, exc, hoff, hits)
    assert n == len(whits) and np.array_equal(hoff, woff) and np.array_equal(hits[:n], whits)
    hoff2, hits2 = np.zeros(len(qo), np.uint64), np.zeros((len(whits), 2), np.uint64)       # pageable outputs
    assert dna_dev.locate_prepacked_into(crumbs, qo2, exc, hoff2, hits2, sorted_hits=True) == len(whits)
    woff2, whits2, _ = dna_or.locate_batch(qb, qo, sorted_hits=True)
    assert np.array_equal(hoff2, woff2) and np.array_equal(hits2, whits2)


def test_prepacked_reads_many_chunks_and_bad_input(dna, dna_dev, dna_or, monkeypatch):
    from awry_b200 import AwryError, fm_index as f
    monkeypatch.setenv("AWRY_B200_LOCATE_CHUNK_Q", "1500")
    qb, qo = ragged_queries(dna, 9_000, seed=2)
    crumbs, exc = f.host_pack_dna(qb)
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    hoff, hits = pinned(len(qo)), pinned((len(whits), 2))
    assert dna_dev.locate_prepacked_into(crumbs, qo, exc, hoff, hits) == len(whits)
    assert np.array_equal(hoff, woff) and np.array_equal(hits, whits)
    # a sentinel travels as an exception and is refused like in the ASCII path
    bad = qb.copy()
    at = int(qo[4321]) + 1 if qo[4322] - qo[4321] > 1 else int(qo[4321])
    bad[at] = ord("$")
    c2, e2 = f.host_pack_dna(bad)
    with pytest.raises(AwryError) as e:
        dna_dev.count_prepacked(c2, qo, e2)
    assert e.value.code == -5 and "query 4321 " in str(e.value)
    with pytest.raises(AwryError) as e:                        # unsorted exception list
        dna_dev.count_prepacked(crumbs, qo, exc[::-1].copy())
    assert e.value.code == -1
    assert np.array_equal(dna_dev.count_prepacked(crumbs, qo, exc), dna_or.count_batch(qb, qo)[0])


def test_prepacked_refused_for_amino(fx):
    from awry_b200 import AwryError
    text = fx.gen_text(1, 5000, 6)
    parts = fx.build_parts(text, 1, ratio=4, kmer_len=2)
    with device_from_parts(parts) as ix:
        with pytest.raises(AwryError) as e:
            ix.count_prepacked(np.zeros(4, np.uint8), np.array([0, 8], np.uint64))
        assert e.value.code == -6


@pytest.mark.parametrize("host_pack", [0, 1])
def test_non_monotone_offsets_are_refused(fx, dna, dna_dev, dna_or, host_pack):
    """an interior offset beyond the batch used to give a query a length of 10^6 and run the pack
    kernel past both buffers.  Now: AWRY_ERR_INVALID_ARG, nothing is loaded or stored for that query, and
    the device is still usable."""
    import torch
    from awry_b200 import AwryError, fm_index as f
    f.set_host_pack(host_pack)
    try:
        qb, qo = mixed_queries(fx, dna.text, 6000, 40, seed=3)
        want, _ = dna_or.count_batch(qb, qo)
        for bad_at, val in ((1, 1_000_000), (3000, 0), (5999, 10**12), (17, int(qo[16]) - 1)):
            o = qo.copy()
            o[bad_at] = val
            with pytest.raises(AwryError) as e:
                dna_dev.count_packed(qb, o)
            assert e.value.code == -1 and "monotone" in str(e.value)
            with pytest.raises(AwryError) as e:
                dna_dev.locate_packed(qb, o)
            assert e.value.code == -1
            hoff, hits = pinned(len(o)), pinned((20000, 2))
            with pytest.raises(AwryError) as e:
                dna_dev.locate_packed_into(qb, o, hoff, hits)
            assert e.value.code == -1
        o = qo.copy()
        o[-1] = 5                                           # the ends themselves
        with pytest.raises(AwryError) as e:
            dna_dev.count_packed(qb, o)
        assert e.value.code == -1
        # device-resident entry point: latched, reported by device_check
        d_qb = torch.from_numpy(qb).cuda()
        o = qo.copy()
        o[100] = 10**9
        d_qo = torch.from_numpy(o.astype(np.int64)).cuda()
        d_c = torch.zeros(6000, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        dna_dev.count_device(d_qb.data_ptr(), d_qo.data_ptr(), 6000, d_c.data_ptr(), st)
        with pytest.raises(AwryError) as e:
            dna_dev.device_check(st)
        assert e.value.code == -1
        torch.cuda.synchronize()
        assert np.array_equal(dna_dev.count_packed(qb, qo), want)          # no sticky error, no corruption
    finally:
        f.set_host_pack(-1)


def test_direct_locate_capacity_and_order(fx, dna, dna_dev, dna_or, monkeypatch):
    """the sync-free pass 2 (pinned hit buffer): exact fit, one short, zero; results equal to the two-sync
    pipeline (AWRY_B200_LOCATE_DIRECT=0) whatever the chunking"""
    from awry_b200 import AwryError
    qb, qo = mixed_queries(fx, dna.text, 30_000, 11, seed=9)          # 11-mers on 300 kbp: 0..few hits
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    n = len(whits)
    assert n > 10_000
    for chunk_q in (1024, 7000, 1 << 19):
        monkeypatch.setenv("AWRY_B200_LOCATE_CHUNK_Q", str(chunk_q))
        hoff, hits = pinned(len(qo)), pinned((n, 2))
        assert dna_dev.locate_packed_into(qb, qo, hoff, hits) == n
        assert np.array_equal(hoff, woff) and np.array_equal(hits, whits)
        short = pinned((n - 1, 2))
        short[:] = 0xAB
        with pytest.raises(AwryError) as e:
            dna_dev.locate_packed_into(qb, qo, hoff, short)
        assert e.value.code == -8 and e.value.needed == n
        assert np.array_equal(hoff, woff) and np.array_equal(short, whits[:n - 1])   # complete up to the capacity
    monkeypatch.setenv("AWRY_B200_LOCATE_DIRECT", "0")
    hoff, hits = pinned(len(qo)), pinned((n, 2))
    assert dna_dev.locate_packed_into(qb, qo, hoff, hits) == n
    assert np.array_equal(hoff, woff) and np.array_equal(hits, whits)


def test_one_batch_over_several_replicas(fx, dna, dna_or):
    """fm_index.rs:455-487 is ONE call over ONE batch: with several devices behind the handle the batch is
    split by query bytes, every replica thread packs through the shared host pool, results come back in
    input order.  ASCII and pre-packed input, count / locate / locate_into."""
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    from awry_b200 import fm_index as f
    devs = list(range(min(n_dev, 8)))
    qb, qo = ragged_queries(dna, 120_001, seed=4)
    want, _ = dna_or.count_batch(qb, qo)
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    crumbs, exc = f.host_pack_dna(qb)
    pq, po_ = pinned(len(qb), np.uint8), pinned(len(qo))
    pq[:] = qb
    po_[:] = qo
    f.set_host_threads(32)
    try:
        with device_from_parts(dna, devices=devs) as ix:
            assert ix.n_devices() == len(devs)
            for mode in (1, 0, -1):
                f.set_host_pack(mode)
                assert np.array_equal(ix.count_packed(pq, po_), want)
                assert np.array_equal(ix.count_packed(qb, qo), want)
            assert np.array_equal(ix.count_prepacked(crumbs, qo, exc), want)
            off, hits = ix.locate_packed(qb, qo)
            assert np.array_equal(off, woff) and np.array_equal(hits, whits)
            hoff, hbuf = pinned(len(qo)), pinned((len(whits), 2))
            assert ix.locate_packed_into(pq, po_, hoff, hbuf) == len(whits)
            assert np.array_equal(hoff, woff) and np.array_equal(hbuf, whits)
            assert ix.locate_prepacked_into(crumbs, qo, exc, hoff, hbuf) == len(whits)
            assert np.array_equal(hoff, woff) and np.array_equal(hbuf, whits)
    finally:
        f.set_host_pack(-1)
        f.set_host_threads(0)


@pytest.mark.parametrize("ratio,lean", [(2, None), (7, None), (8, None), (32, None), (64, 64), (8, 5), (32, 33), (3, 16)])
def test_bounded_walk_is_the_default_without_the_unsampled_array(fx, po, monkeypatch, ratio, lean):
    """without the 4-byte-per-row array (AWRY_B200_FULL_SA=0) locate uses the suffix array
    sampled by text position -- every walk ends within ratio - 1 steps -- and returns what the reference's
    row-sampled walk returns, hit for hit, in BWT-row order.  Multi-record text with N runs included."""
    from awry_b200 import fm_index as f
    monkeypatch.setenv("AWRY_B200_FULL_SA", "0")
    monkeypatch.delenv("AWRY_B200_LEAN_SA", raising=False)
    if lean:
        monkeypatch.setenv("AWRY_B200_LEAN_RATIO", str(lean))   # the sampling distance is the library's, not the file's
    recs = [bytes(fx.gen_text(0, n, 40 + i)) for i, n in enumerate([5000, 223, 224, 225, 1, 30_000, 447])]
    recs[5] = recs[5][:1000] + b"NNNNNNNNNN" + recs[5][1010:]
    text, starts = fx.concat_records(recs, 0)
    parts = fx.build_parts(text, 0, ratio=ratio, kmer_len=5, seq_starts=starts, headers=[f"r{i}" for i in range(len(recs))])
    orc = oracle_from_parts(po, parts)
    qs = []
    for r in recs:
        qs += [r[i:i + 9] for i in range(0, max(1, len(r) - 9), 97)] + [r[:3], r[-2:]]
    qs += [b"A", b"C", b"N", b"NN", b"AN", b"TTTT", b"ACG"]
    qb, qo = f.pack_queries(qs)
    woff, whits, _ = orc.locate_batch(qb, qo)
    with device_from_parts(parts) as ix:
        assert ix.device_bytes()["full_sa"] == 0 and ix.device_bytes()["lean_sa"] > 0
        assert ix.lean_sa_ratio() == (lean or min(ratio, 4))
        off, hits = ix.locate_packed(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        soff, shits = ix.locate_packed(qb, qo, sorted_hits=True)
        woff2, whits2, _ = orc.locate_batch(qb, qo, sorted_hits=True)
        assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)
        f.set_locate_variant(1)                                  # the reference's own scheme on the same handle
        try:
            off1, hits1 = ix.locate_packed(qb, qo)
        finally:
            f.set_locate_variant(0)
        assert np.array_equal(off1, woff) and np.array_equal(hits1, whits)


def test_corrupt_index_files_fail_with_format_errors(tmp_path):
    """untrusted file fields -- k-mer length, prefix sums, sequence starts and count -- are validated;
    a corrupt file is AWRY_ERR_FORMAT, never an out-of-range row pointer or a load that does not end"""
    import os
    from awry_b200 import AwryError, FmIndex
    golden = open(os.path.join(os.path.dirname(__file__), "golden", "appendix_a.awry"), "rb").read()
    # layout of the 557-byte file: blocks @43 (160 B), prefix sums @203 (7 x 8), SA @259 (8 B),
    # k byte @267, table 16 x 16 B, sequence index @524: n, start, header_len, header
    def variant(name, edit):
        b = bytearray(golden)
        edit(b)
        p = tmp_path / f"{name}.awry"
        p.write_bytes(bytes(b))
        return str(p)

    def put(b, at, v):
        b[at:at + 8] = int(v).to_bytes(8, "little")

    cases = {
        "k_huge": lambda b: b.__setitem__(267, 40),
        "prefix_decreasing": lambda b: (put(b, 203 + 16, 12), put(b, 203 + 24, 9)),
        "prefix_first": lambda b: put(b, 203 + 8, 2),
        "prefix_total": lambda b: put(b, 203 + 48, 21),
        "seq_count": lambda b: put(b, 524, 2**32),
        "seq_start_beyond": lambda b: put(b, 532, 20),
        "alphabet": lambda b: put(b, 35, 2),
        "ratio_zero": lambda b: put(b, 19, 0),
    }
    for name, edit in cases.items():
        with pytest.raises(AwryError) as e:
            FmIndex.load(variant(name, edit))
        assert e.value.code == -3, (name, str(e.value))
    with FmIndex.load(variant("intact", lambda b: None)) as ix:
        assert ix.count_string("GATTACA") == 2
