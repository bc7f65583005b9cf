"""Wide indexes: 64-bit row pointers (SearchPtr = u64, /root/reference/src/search.rs:7).  The wide path is
taken from bwt_len = 2^32 - 256 on; AWRY_B200_WIDE=1 forces it on any index and AWRY_B200_SB_SHIFT shrinks the
superblocks the block counts are relative to, so the small indexes below cross many superblock borders.  Same
bar as everywhere: bit-exact versus the oracle.  The last test builds, loads and searches a real 4.6 Gbp index."""
import os

import numpy as np
import pytest

from conftest import brute_positions, device_from_parts, mixed_queries, oracle_from_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[31, 12, 8], ids=["sb2^31", "sb4096", "sb256"])
def wide_env(request, monkeypatch):
    monkeypatch.setenv("AWRY_B200_WIDE", "1")
    monkeypatch.setenv("AWRY_B200_SB_SHIFT", str(request.param))
    return request.param


def test_wide_dna_matches_oracle(fx, po, wide_env, tmp_path):
    from awry_b200 import FmIndex, SearchRange, fm_index as f
    recs = [bytes(fx.gen_text(0, n, 60 + i)) for i, n in enumerate([40_000, 300, 1, 9_000])]
    recs[0] = recs[0][:500] + b"NNNNRYK" + recs[0][507:]
    text, starts = fx.concat_records(recs, 0)
    parts = fx.build_parts(text, 0, ratio=5, kmer_len=6, seq_starts=starts, headers=[f"r{i}" for i in range(len(recs))])
    orc = oracle_from_parts(po, parts)
    qb, qo = mixed_queries(fx, text, 6000, 21, seed=2)
    t = bytes(text)
    edge = [t[5:5 + n] for n in range(1, 9)] + [t[100:140].lower(), b"N", b"NN", b"ACGTNACGT", t[:300], t[-40:], b"A" * 40,
                                                 t + b"A", b"acgtn", t[490:520]]
    eb, eo = f.pack_queries(edge)
    with device_from_parts(parts) as ix:
        assert ix.row_pointer_bits() == 64
        assert ix.device_bytes()["pair"] == 0 and ix.device_bytes()["full_sa"] == 0 and ix.device_bytes()["lean_sa"] == 0
        for b, o in ((qb, qo), (eb, eo)):
            want, _ = orc.count_batch(b, o)
            assert np.array_equal(ix.count_packed(b, o), want)           # cooperative kernel (4 lanes per query)
            f.set_search_variant(-1)                                     # one thread per query
            try:
                assert np.array_equal(ix.count_packed(b, o), want)
            finally:
                f.set_search_variant(0)
            off, hits = ix.locate_packed(b, o)
            woff, whits, _ = orc.locate_batch(b, o)
            assert np.array_equal(off, woff) and np.array_equal(hits, whits)
            soff, shits = ix.locate_packed(b, o, sorted_hits=True)
            woff2, whits2, _ = orc.locate_batch(b, o, sorted_hits=True)
            assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)
        for q, r in zip(edge, ix.search_packed(eb, eo)):
            sp, ep = orc.search_range(q)
            assert (int(r[0]), int(r[1])) == ((sp, ep) if sp <= ep else (1, 0)), q
        rng = ix.initial_search_range("G")
        assert tuple(rng) == orc.initial_range(orc.sym("G"))
        for ch in "ACGTNacgu":
            assert tuple(ix.update_range_with_symbol(rng, ch)) == orc.update_range(rng.start_ptr, rng.end_ptr, orc.sym(ch))
        for row in [0, 1, 127, 128, 255, 256, 4095, 4096, parts.bwt_len - 1] + [int(x) for x in np.random.default_rng(0).integers(0, parts.bwt_len, 40)]:
            assert ix.backstep(row) == orc.backstep(row), row
        assert ix.initial_search_range("$") == SearchRange(0, 0)
        # save -> byte-identical to the reference-layout writer; load of that file -> same answers
        ref_file, out_file = str(tmp_path / "ref.awry"), str(tmp_path / "out.awry")
        parts.write(ref_file)
        ix.save(out_file)
        assert open(ref_file, "rb").read() == open(out_file, "rb").read()
    with FmIndex.load(ref_file) as ix2:
        assert ix2.row_pointer_bits() == 64 and ix2.sequence_header(3) == "r3"
        assert np.array_equal(ix2.count_packed(qb, qo), orc.count_batch(qb, qo)[0])
        off, hits = ix2.locate_packed(eb, eo)
        woff, whits, _ = orc.locate_batch(eb, eo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    for i in (0, 5, 3000):
        q = bytes(qb[21 * i:21 * i + 21])
        assert int(orc.count_batch(qb, qo)[0][i]) == len(brute_positions(t, q))


def test_wide_amino_matches_oracle(fx, po, wide_env, tmp_path):
    from awry_b200 import fm_index as f
    text = fx.gen_text(1, 30_000, 6)
    parts = fx.build_parts(text, 1, ratio=8, kmer_len=3)
    orc = oracle_from_parts(po, parts)
    qb, qo = mixed_queries(fx, text, 3000, 4, seed=7, alphabet=1)
    qs = [bytes(qb[4 * i:4 * i + 4]) for i in range(3000)]
    qs += [b"A", b"AC", b"X", b"acd", b"BZJ", b"W" * 9, bytes(text[:300]), bytes(text[-20:]), b"M*K", b"ACDEFGHIKLMNPQRSTVWY"]
    qb, qo = f.pack_queries(qs)
    with device_from_parts(parts) as ix:
        assert ix.row_pointer_bits() == 64
        assert np.array_equal(ix.count_packed(qb, qo), orc.count_batch(qb, qo)[0])
        off, hits = ix.locate_packed(qb, qo)
        woff, whits, _ = orc.locate_batch(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        for row in (0, 1, 63, 64, 65, 12345, parts.bwt_len - 1):
            assert ix.backstep(row) == orc.backstep(row)
        out_file, ref_file = str(tmp_path / "o.awry"), str(tmp_path / "r.awry")
        parts.write(ref_file)
        ix.save(out_file)
        assert open(ref_file, "rb").read() == open(out_file, "rb").read()


@pytest.mark.parametrize("alphabet", [0, 1])
def test_wide_builder_equals_cpu_builder(fx, po, monkeypatch, alphabet):
    """the bucket-by-bucket suffix sort with 64-bit positions (texts of 2^32 symbols or more), forced on small
    texts: the same reference-layout arrays as the CPU builder, bit for bit -- repeats and ties included"""
    from awry_b200 import fm_index as f
    monkeypatch.setenv("AWRY_B200_BUILD_WIDE", "1")
    r = np.random.default_rng(5)
    base = fx.gen_text(alphabet, 50_000, 9)
    text = np.concatenate([base, base[1000:1300], base[:700], fx.gen_text(alphabet, 3000, 10), base[1000:1300]])
    if alphabet == 0:
        text = text.copy()
        text[20_000:20_006] = np.frombuffer(b"NNRYKN", dtype=np.uint8)
    for ratio in (1, 8, 13):
        want = fx.build_parts(text, alphabet, ratio=ratio, kmer_len=4)
        blocks, prefix, sa_words, _ = f.build_parts(alphabet, text, sa_ratio=ratio)
        assert np.array_equal(prefix, want.prefix_sums)
        assert np.array_equal(blocks, want.blocks)
        assert np.array_equal(sa_words, want.sa_words)


def test_wide_index_4p6_gbp_builds_loads_counts_locates(fx, po):
    """a real wide index: 4.6 G rows (> 2^32), built on the GPU with 64-bit suffix-array elements, handed over
    in the reference layout, counted and located bit-exact versus the oracle on a sample; every exact-substring
    read found; located positions hold the query (text regenerated from its seed)"""
    import torch
    from awry_b200 import FmIndex
    from fixtures import pyfixture_gpu as fxg
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 110e9:
        pytest.skip(f"needs ~100 GB of free HBM for the 64-bit suffix sort ({free / 1e9:.0f} GB free)")
    n, seed, k, ratio = 4_600_000_000, 12, 12, 16
    parts, phases = fxg.build_parts(0, n, seed, ratio=ratio, kmer_len=k)
    assert parts.bwt_len == n + 1 > 2**32
    orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                    parts.prefix_sums, parts.sa_words)
    st = torch.cuda.current_stream().cuda_stream
    with FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words) as ix:
        assert ix.row_pointer_bits() == 64 and ix.bwt_len() == n + 1
        assert ix.prefix_sums()[-1] == n + 1
        nq, L = 2_000_000, 100
        d_q = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, n, seed, nq, L, 4, d_q.data_ptr())
        d_off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * L
        d_cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), nq, d_cnt.data_ptr(), st)
        ix.device_check(st)
        assert int(d_cnt.min()) >= 1
        ns = 20_000
        qb = d_q[: ns * L].cpu().numpy()
        qo = np.arange(ns + 1, dtype=np.uint64) * np.uint64(L)
        want, _ = orc.count_batch(qb, qo)
        assert np.array_equal(want, d_cnt[:ns].cpu().numpy().view(np.uint64))
        assert np.array_equal(ix.count_packed(qb, qo), want)
        # locate: short queries (several hits), positions beyond 2^32 among them
        nl, ll = 20_000, 17
        lq = d_q[: nl * L].cpu().numpy().reshape(nl, L)[:, :ll].copy().reshape(-1)
        lo = np.arange(nl + 1, dtype=np.uint64) * np.uint64(ll)
        off, hits = ix.locate_packed(lq, lo)
        woff, whits, _ = orc.locate_batch(lq, lo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        assert int(hits[:, 1].max()) > 2**32
        sel = np.arange(0, len(hits), max(1, len(hits) // 5000))
        qidx = np.searchsorted(off, sel, side="right") - 1
        windows = fx.gen_text_windows(0, seed, hits[sel, 1], ll)
        assert np.array_equal(windows, lq.reshape(nl, ll)[qidx])
