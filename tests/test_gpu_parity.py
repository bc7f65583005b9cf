"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bar: bit-exact counts, ranges and locate lists (integer work)."""
import os

import numpy as np
import pytest

from conftest import (brute_positions, device_from_parts, mixed_queries, oracle_from_parts)

pytestmark = pytest.mark.gpu

APPENDIX_A = {
    "A": ((1, 7), [4, 11, 15, 6, 13, 1, 8]),
    "TTA": ((18, 19), [2, 9]),
    "GATTACA": ((11, 12), [0, 7]),
    "ACA": ((1, 2), [4, 11]),
    "N": ((14, 14), [14]),
    "CAN": ((9, 9), [12]),
    "ACGT": ((3, 3), [15]),
    "GG": ((1, 0), []),
    "TACAGATTACANACGT": ((16, 16), [3]),
}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "appendix_a.awry")


@pytest.fixture(scope="module")
def dna(fx):
    text = fx.gen_text(0, 200_000, 1)
    return fx.build_parts(text, 0, ratio=8, kmer_len=8)


@pytest.fixture(scope="module")
def dna_dev(dna):
    ix = device_from_parts(dna)
    yield ix
    ix.close()


@pytest.fixture(scope="module")
def dna_or(po, dna):
    return oracle_from_parts(po, dna)


@pytest.fixture(autouse=True, params=[1, 0], ids=["hostpack", "ascii"])
def host_pack_mode(request):
    """every test of this module runs with nucleotide queries packed to 2 bits on the host (the default on
    a many-core host) and with the ASCII bytes sent as they are"""
    from awry_b200 import fm_index as f
    f.set_host_pack(request.param)
    yield request.param
    f.set_host_pack(-1)


@pytest.fixture(params=[0, 1, 2], ids=["unsampled_sa", "lf_walk", "bounded_walk"])
def locate_variant(request):
    """locate pass 2 all three ways: gather from the unsampled suffix array (default when it fits), the
    LF-walk to the sampled rows (fm_index.rs:521-537), and the bounded walk on the position-sampled array"""
    from awry_b200 import fm_index as f
    f.set_locate_variant(request.param)
    yield request.param
    f.set_locate_variant(0)


def edge_queries(text):
    t = bytes(text[:400])
    qs = [t[5:5 + n] for n in range(1, 14)]                 # shorter than / equal to / just above k
    qs += [t[40:72].lower(), t[100:130].replace(b"T", b"U"), b"acgtn", b"N", b"NN", b"ACGTNACGT",
           t[10:30] + b"N" + t[31:50], b"R", b"x", t[0:300], bytes(text[-40:]), bytes(text[-1:]),
           b"A" * 64, b"ACGT" * 40, bytes(text) + b"A", b"TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTG"]
    return qs


def test_golden_file_appendix_a(locate_variant):
    from awry_b200 import FmIndex, SearchRange
    with FmIndex.load(GOLDEN) as ix:
        assert ix.bwt_len() == 20 and ix.suffix_array_compression_ratio() == 4
        assert ix.prefix_sums() == [0, 1, 8, 11, 14, 15, 20]
        assert ix.kmer_len() == 2 and ix.sequence_header(0) == "synthetic"
        for q, (rng, locs) in APPENDIX_A.items():
            assert ix.get_search_range_for_string(q) == SearchRange(*rng), q
            assert ix.count_string(q) == len(locs), q
            assert [h.local_position for h in ix.locate_string(q)] == locs, q
            assert all(h.sequence_idx == 0 for h in ix.locate_string(q))


@pytest.mark.parametrize("lanes", [0, 1, 2, 4, 8, -1])
def test_count_parity_cfg1_style(fx, dna, dna_dev, dna_or, lanes):
    """parallel_count of 10k random 32-bp queries (half present, half random) + edge cases."""
    from awry_b200 import fm_index as f
    qb, qo = mixed_queries(fx, dna.text, 10_000, 32, seed=2)
    f.set_search_variant(lanes)
    try:
        got = dna_dev.count_packed(qb, qo)
        want, _ = dna_or.count_batch(qb, qo)
        assert np.array_equal(got, want)
        assert int((want > 0).sum()) >= 5000
        qs = edge_queries(dna.text)
        eb, eo = f.pack_queries(qs)
        got = dna_dev.count_packed(eb, eo)
        want, _ = dna_or.count_batch(eb, eo)
        assert np.array_equal(got, want), [(q[:20], int(a), int(b)) for q, a, b in zip(qs, got, want) if a != b]
        rg = dna_dev.search_packed(eb, eo)
        for q, r in zip(qs, rg):
            sp, ep = dna_or.search_range(q)
            assert (int(r[0]), int(r[1])) == ((sp, ep) if sp <= ep else (1, 0)), q
    finally:
        f.set_search_variant(0)


def test_counts_match_brute_force(fx, dna, dna_dev):
    text = bytes(dna.text)
    qb, qo, _ = fx.gen_substring_queries(dna.text, 300, 11, seed=5)
    got = dna_dev.count_packed(qb, qo)
    for i in range(300):
        q = bytes(qb[11 * i:11 * i + 11])
        assert int(got[i]) == len(brute_positions(text, q))


@pytest.mark.parametrize("ratio", [1, 4, 5, 8, 32])
def test_locate_parity(fx, po, ratio, locate_variant):
    text = fx.gen_text(0, 60_000, 11 + ratio)
    parts = fx.build_parts(text, 0, ratio=ratio, kmer_len=6)
    orc = oracle_from_parts(po, parts)
    qb1, qo1, _ = fx.gen_substring_queries(text, 2000, 20, seed=3)   # mostly unique
    qb2, qo2, _ = fx.gen_substring_queries(text, 500, 5, seed=4)     # ~60 hits each
    from awry_b200 import fm_index as f
    qs = [bytes(qb1[20 * i:20 * i + 20]) for i in range(2000)] + [bytes(qb2[5 * i:5 * i + 5]) for i in range(500)]
    qs += [b"A", b"ACGTTTTTTTTGGGGGGGGGGGGGCCCCCCCCCCC", bytes(text[:50]), bytes(text[-30:])]
    qb, qo = f.pack_queries(qs)
    with device_from_parts(parts) as ix:
        assert ix.device_bytes()["full_sa"] == (4 * parts.bwt_len if ratio > 1 else 0)
        assert (ix.device_bytes()["lean_sa"] > 0) == (ratio > 1)
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, _ = orc.locate_batch(qb, qo)
        assert np.array_equal(off, woff)
        assert np.array_equal(hits, whits)          # BWT-row order == the reference's push order
        soff, shits = ix.locate_packed(qb, qo, sorted_hits=True)
        woff2, whits2, _ = orc.locate_batch(qb, qo, sorted_hits=True)
        assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)
    # oracle-of-the-oracle on a few
    t = bytes(text)
    for i in (0, 1, 2000, 2001, 2500):
        lo, hi = int(woff[i]), int(woff[i + 1])
        assert sorted(int(x) for x in whits[lo:hi, 1]) == brute_positions(t, qs[i])


def test_amino_count_and_locate(fx, po, locate_variant):
    text = fx.gen_text(1, 50_000, 6)
    parts = fx.build_parts(text, 1, ratio=8, kmer_len=3)
    orc = oracle_from_parts(po, parts)
    from awry_b200 import fm_index as f
    qb, qo = mixed_queries(fx, text, 4000, 4, seed=7, alphabet=1)
    qs = [bytes(qb[4 * i:4 * i + 4]) for i in range(4000)]
    qb3, qo3, _ = fx.gen_substring_queries(text, 1000, 12, seed=8)
    qs += [bytes(qb3[12 * i:12 * i + 12]) for i in range(1000)]
    qs += [b"A", b"AC", b"X", b"acd", b"BZJ", b"W" * 9, bytes(text[:300]), bytes(text[-20:]), b"M*K", b"ACDEFGHIKLMNPQRSTVWY"]
    qb, qo = f.pack_queries(qs)
    with device_from_parts(parts) as ix:
        want, _ = orc.count_batch(qb, qo)
        for variant in (0, -1):          # cooperative 4-lane kernel, scalar kernel
            f.set_search_variant(variant)
            try:
                got = ix.count_packed(qb, qo)
            finally:
                f.set_search_variant(0)
            assert np.array_equal(got, want), variant
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, _ = orc.locate_batch(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        for row in (0, 1, 63, 64, 65, 12345, parts.bwt_len - 1):
            assert ix.backstep(row) == orc.backstep(row)
    t = bytes(text)
    for i in (0, 5, 4000, 4500):
        assert int(want[i]) == len(brute_positions(t, qs[i]))


def test_multi_record_and_n_text(fx, po, locate_variant):
    """multi-record input: delimiters are literal N symbols (fm_index.rs:148-153); locate maps
    to (record, offset) with the intended semantics (reference recursion: SURVEY Q4)."""
    recs = [bytes(fx.gen_text(0, n, 20 + i)) for i, n in enumerate([700, 33, 1, 1500, 256, 90])]
    recs[3] = recs[3][:100] + b"NNNNRYK" + recs[3][107:]
    text, starts = fx.concat_records(recs, 0)
    parts = fx.build_parts(text, 0, ratio=3, kmer_len=4, seq_starts=starts,
                           headers=[f"rec{i}" for i in range(len(recs))])
    orc = oracle_from_parts(po, parts)
    from awry_b200 import fm_index as f
    qs = []
    for r in recs:
        qs += [r[i:] for i in range(0, len(r), 37)] + [r[:5], r[-3:]]
    qs += [b"N", b"NN", b"NNNN", b"AN", b"NA", b"ACGN", b"NNNNN"]
    qb, qo = f.pack_queries(qs)
    with device_from_parts(parts) as ix:
        assert ix.sequence_header(3) == "rec3"
        got = ix.count_packed(qb, qo)
        want, _ = orc.count_batch(qb, qo)
        assert np.array_equal(got, want)
        assert all(int(c) > 0 for c in got[:len(qs) - 7])     # fm_index.rs:779-790 property
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, _ = orc.locate_batch(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        assert int(hits[:, 0].max()) == len(recs) - 1


def test_invalid_queries_are_errors(dna_dev):
    from awry_b200 import AwryError, fm_index as f
    for bad in ([b"ACGT", b""], [b"AC$T"], [b"#"], [b"ACGT", b"GGA$"]):
        qb, qo = f.pack_queries(bad)
        with pytest.raises(AwryError) as e:
            dna_dev.count_packed(qb, qo)
        assert e.value.code == -5
        with pytest.raises(AwryError):
            dna_dev.locate_packed(qb, qo)
    assert dna_dev.count_string("ACGT") >= 0   # index still usable afterwards


def test_invalid_and_ambiguous_queries_in_large_batches(fx, dna, dna_dev, dna_or):
    """batches large enough for the host-packed path: sentinels and empty queries are still reported by
    index, N / IUPAC / lower-case bytes travel as exceptions (few) or force the ASCII path (many)"""
    from awry_b200 import AwryError, fm_index as f
    qb, qo, _ = fx.gen_substring_queries(dna.text, 3000, 50, seed=21)
    qs = [bytes(qb[50 * i:50 * i + 50]) for i in range(3000)]
    for i in range(0, 3000, 97):                      # ~1 % ambiguous bytes: the exception list
        q = bytearray(qs[i])
        q[i % 50] = ord("N")
        q[(i * 7) % 50] = ord("r")
        qs[i] = bytes(q)
    qs[5] = qs[5].lower()
    qs[6] = qs[6].replace(b"T", b"U")
    b, o = f.pack_queries(qs)
    want, _ = dna_or.count_batch(b, o)
    assert np.array_equal(dna_dev.count_packed(b, o), want)
    woff, whits, _ = dna_or.locate_batch(b, o)
    off, hits = dna_dev.locate_packed(b, o)
    assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    many_n = [q[:10] + b"NNNNNNNN" + q[18:] for q in qs]   # 16 % ambiguous: the chunk goes up as ASCII
    b2, o2 = f.pack_queries(many_n)
    want2, _ = dna_or.count_batch(b2, o2)
    assert np.array_equal(dna_dev.count_packed(b2, o2), want2)
    for bad_at, bad in ((1234, qs[1234][:20] + b"$" + qs[1234][21:]), (2999, b""), (0, b"#" + qs[0][1:])):
        bad_qs = list(qs)
        bad_qs[bad_at] = bad
        b3, o3 = f.pack_queries(bad_qs)
        with pytest.raises(AwryError) as e:
            dna_dev.count_packed(b3, o3)
        assert e.value.code == -5 and f"query {bad_at} " in str(e.value)
    assert np.array_equal(dna_dev.count_packed(b, o), want)


def test_single_step_api(dna, dna_dev, dna_or):
    from awry_b200 import SearchRange
    rng = dna_dev.initial_search_range("G")
    assert tuple(rng) == dna_or.initial_range(dna_or.sym("G"))
    for ch in "ACGTNacgu":
        nxt = dna_dev.update_range_with_symbol(rng, ch)
        assert tuple(nxt) == dna_or.update_range(rng.start_ptr, rng.end_ptr, dna_or.sym(ch)), ch
    r = np.random.default_rng(0)
    for row in [0, 1, 127, 128, 255, 256, dna.bwt_len - 1] + [int(x) for x in r.integers(0, dna.bwt_len, 40)]:
        assert dna_dev.backstep(row) == dna_or.backstep(row), row
    assert dna_dev.initial_search_range("$") == SearchRange(0, 0)


def test_file_load_equals_parts(tmp_path, fx, dna, dna_dev, dna_or, po):
    from awry_b200 import FmIndex
    path = str(tmp_path / "t.awry")
    dna.write(path)
    qb, qo = mixed_queries(fx, dna.text, 3000, 24, seed=12)
    with FmIndex.load(path) as ix:
        assert ix.prefix_sums() == [int(x) for x in dna.prefix_sums]
        assert ix.bwt_len() == dna.bwt_len and ix.kmer_len() == 8
        assert np.array_equal(ix.count_packed(qb, qo), dna_dev.count_packed(qb, qo))
        a = ix.locate_packed(qb, qo)
        b = dna_dev.locate_packed(qb, qo)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    orc = po.OracleIndex.load(path)
    assert np.array_equal(orc.count_batch(qb, qo)[0], dna_dev.count_packed(qb, qo))


def test_device_resident_entry_points(fx, dna, dna_dev, dna_or):
    import torch
    qb, qo = mixed_queries(fx, dna.text, 20_000, 40, seed=14)
    d_qb = torch.from_numpy(qb).cuda()
    d_qo = torch.from_numpy(qo.astype(np.int64)).cuda()
    d_counts = torch.zeros(20_000, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    dna_dev.count_device(d_qb.data_ptr(), d_qo.data_ptr(), 20_000, d_counts.data_ptr(), st)
    dna_dev.device_check(st)
    want, _ = dna_or.count_batch(qb, qo)
    assert np.array_equal(d_counts.cpu().numpy().astype(np.uint64), want)
    d_off = torch.zeros(20_001, dtype=torch.int64, device="cuda")
    ptr, n = dna_dev.locate_device(d_qb.data_ptr(), d_qo.data_ptr(), 20_000, d_off.data_ptr(), stream=st)
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    assert n == len(whits)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), woff)
    dna_dev.device_free(ptr)


def test_ragged_lengths_and_order_preserved(fx, dna, dna_dev, dna_or):
    from awry_b200 import fm_index as f
    r = np.random.default_rng(3)
    t = bytes(dna.text)
    qs = []
    for _ in range(5000):
        n = int(r.integers(1, 200))
        p = int(r.integers(0, len(t) - n))
        q = bytearray(t[p:p + n])
        if r.random() < 0.3:
            q[int(r.integers(0, n))] = ord("ACGTN"[int(r.integers(0, 5))])
        qs.append(bytes(q))
    qb, qo = f.pack_queries(qs)
    got = dna_dev.count_packed(qb, qo)
    want, _ = dna_or.count_batch(qb, qo)
    assert np.array_equal(got, want)
    assert dna_dev.parallel_count(qs[:50]) == [int(x) for x in want[:50]]
    loc = dna_dev.parallel_locate(qs[:50])
    for i in range(50):
        assert [tuple(h) for h in loc[i]] == dna_or.locate_string(qs[i])


def test_cxx_host_mirror_every_kmer(tmp_path, fx):
    """the C++ mirror of the Rust API (include/awry_b200.hpp), driven like fm_index.rs:612-664"""
    import subprocess
    from conftest import ROOT
    text = fx.gen_text(0, 1847, 0)
    parts = fx.build_parts(text, 0)
    idx, txt, exe = str(tmp_path / "n.awry"), str(tmp_path / "n.txt"), str(tmp_path / "mirror")
    parts.write(idx)
    open(txt, "wb").write(bytes(text))
    lib = os.path.join(ROOT, "awry_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cxx_host_mirror.cpp"), "-o", exe, "-L", lib,
                           "-lawry_b200", f"-Wl,-rpath,{lib}"])
    out = subprocess.run([exe, idx, txt, "24"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok ")


def test_multi_replica_in_one_process(fx, dna, dna_or):
    """awry_index_from_parts with several devices: replica 0 is cloned over NVLink, batches are split
    by query bytes across replicas (one host thread each), results come back in input order."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    from awry_b200 import fm_index as f
    devs = list(range(min(n, 4)))
    qb, qo = mixed_queries(fx, dna.text, 30_001, 36, seed=31)
    with device_from_parts(dna, devices=devs) as ix:
        assert ix.n_devices() == len(devs)
        got = ix.count_packed(qb, qo)
        want, _ = dna_or.count_batch(qb, qo)
        assert np.array_equal(got, want)
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, _ = dna_or.locate_batch(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        few = f.pack_queries([b"ACGT", b"GGGTTTAA", b"A"])      # fewer queries than 2 x replicas
        assert np.array_equal(ix.count_packed(*few), dna_or.count_batch(*few)[0])


def test_cfg5_repeat_rich_locate_under_skew(fx, po, locate_variant):
    """BASELINE cfg 5 (scaled to what the CPU fixture builder sorts in seconds): repeat-rich DNA,
    SA ratio 32, hit counts from 1 to thousands per query -- exercises the warp-cooperative CSR
    expansion and the lane-refilling walk under load imbalance."""
    from fixtures import repeats
    text, regions = repeats.repeat_rich_text(4_000_000, seed=8)
    parts = fx.build_parts(text, 0, ratio=32, kmer_len=8)
    orc = oracle_from_parts(po, parts)
    qb, qo = repeats.repeat_queries(text, regions, 4000, 50, seed=9)
    with device_from_parts(parts) as ix:
        counts = ix.count_packed(qb, qo)
        want, _ = orc.count_batch(qb, qo)
        assert np.array_equal(counts, want)
        assert int(want.max()) >= 1000 and int(want.min()) == 1 and float(np.median(want)) < want.max() / 10
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, st = orc.locate_batch(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        assert st["walk_steps"] / max(1, st["hits"]) > 20          # mean walk ~ ratio - 1
        soff, shits = ix.locate_packed(qb, qo, sorted_hits=True)
        woff2, whits2, _ = orc.locate_batch(qb, qo, sorted_hits=True)
        assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)
    # round trip on the biggest query: every located position holds the query
    i = int(np.argmax(want))
    q = bytes(qb[50 * i:50 * i + 50])
    t = bytes(text)
    lo, hi = int(woff[i]), int(woff[i + 1])
    assert all(t[int(p):int(p) + 50] == q for p in whits[lo:hi, 1])


def test_locate_into_caller_buffers(fx, dna, dna_dev, dna_or, locate_variant):
    import torch
    from awry_b200 import AwryError
    qb, qo = mixed_queries(fx, dna.text, 5000, 14, seed=41)
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    n = len(whits)
    hoff = torch.zeros(5001, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    hits = torch.zeros((n + 7, 2), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    got = dna_dev.locate_packed_into(qb, qo, hoff, hits)
    assert got == n and np.array_equal(hoff, woff) and np.array_equal(hits[:n], whits)
    small = np.zeros((n // 2, 2), dtype=np.uint64)
    with pytest.raises(AwryError) as e:
        dna_dev.locate_packed_into(qb, qo, hoff, small)
    assert e.value.code == -8 and e.value.needed == n and np.array_equal(hoff, woff)
    with pytest.raises(AwryError) as e:
        dna_dev.locate_packed_into(qb, qo, hoff, np.zeros((0, 2), dtype=np.uint64))
    assert e.value.needed == n


def test_concurrent_batches_from_many_host_threads(fx, dna, dna_dev, dna_or):
    """the index is immutable; *_batch calls may run concurrently (like &self methods called from the
    rayon pool in the reference).  Each call takes a private stream + workspace."""
    import threading
    jobs = []
    for t in range(6):
        qb, qo = mixed_queries(fx, dna.text, 4000 + 500 * t, 20 + 3 * t, seed=100 + t)
        jobs.append((qb, qo, dna_or.count_batch(qb, qo)[0], dna_or.locate_batch(qb, qo)[:2]))
    results, errors = [None] * len(jobs), []

    def work(i):
        try:
            qb, qo = jobs[i][0], jobs[i][1]
            for _ in range(3):
                c = dna_dev.count_packed(qb, qo)
                off, hits = dna_dev.locate_packed(qb, qo)
            results[i] = (c, off, hits)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for (qb, qo, want, (woff, whits)), (c, off, hits) in zip(jobs, results):
        assert np.array_equal(c, want) and np.array_equal(off, woff) and np.array_equal(hits, whits)


@pytest.mark.parametrize("chunk_q", [1024, 3000])
def test_locate_pipeline_many_small_chunks(fx, dna, dna_dev, dna_or, locate_variant, monkeypatch, chunk_q):
    """the 3-deep locate pipeline over many chunks (slot reuse, offsets rebased onto the hits before them,
    growth of the library-owned hit buffer while copies are in flight): results must not depend on the
    chunk size, whether the output is library-owned or caller-owned"""
    import torch
    monkeypatch.setenv("AWRY_B200_LOCATE_CHUNK_Q", str(chunk_q))      # read per call
    qb, qo = mixed_queries(fx, dna.text, 20_000, 12, seed=77)         # 12-mers on 200 kbp: 0 .. several hits each
    woff, whits, _ = dna_or.locate_batch(qb, qo)
    assert int(np.diff(woff).max()) >= 2 and int(np.diff(woff).min()) == 0
    off, hits = dna_dev.locate_packed(qb, qo)
    assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    soff, shits = dna_dev.locate_packed(qb, qo, sorted_hits=True)
    woff2, whits2, _ = dna_or.locate_batch(qb, qo, sorted_hits=True)
    assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)
    n = len(whits)
    hoff = torch.zeros(len(qo), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    hbuf = torch.zeros((n + 3, 2), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    assert dna_dev.locate_packed_into(qb, qo, hoff, hbuf) == n
    assert np.array_equal(hoff, woff) and np.array_equal(hbuf[:n], whits)
