"""BASELINE.json full-size configurations on the GPU, checked through size-independent properties
(the oracle is only run on samples at this size):
  * every exact-substring read has count >= 1, and a sample of counts is bit-exact vs the oracle;
  * locate round trip: the text regenerated at every reported position equals the query;
  * hit totals equal the count totals (a checksum of checksums between the two entry points).
Index fixtures are rebuilt on the GPU in ~2 s per configuration."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device_queries(fxg, alphabet, n, text_seed, nq, qlen, qseed):
    import torch
    d = torch.empty(nq * qlen, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(alphabet, n, text_seed, nq, qlen, qseed, d.data_ptr())
    off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * qlen
    return d, off


def _run_config(fx, po, alphabet, n, k, ratio, count_nq, count_len, loc_nq, loc_len, text_seed, sample=50_000):
    import ctypes as C
    import torch
    from awry_b200 import FmIndex, fm_index as f
    from fixtures import pyfixture_gpu as fxg
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~100 GB of free HBM for the fixture build")
    parts, _ = fxg.build_parts(alphabet, n, text_seed, ratio=ratio, kmer_len=k)
    orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                    parts.prefix_sums, parts.sa_words)
    st = torch.cuda.current_stream().cuda_stream
    with FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words) as ix:
        # ---- count at full size
        d_q, d_off = _device_queries(fxg, alphabet, n, text_seed, count_nq, count_len, 4)
        d_cnt = torch.zeros(count_nq, dtype=torch.int64, device="cuda")
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), count_nq, d_cnt.data_ptr(), st)
        ix.device_check(st)
        assert int(d_cnt.min()) >= 1                       # every read is a substring of the text
        qb = d_q[: sample * count_len].cpu().numpy()
        qo = np.arange(sample + 1, dtype=np.uint64) * np.uint64(count_len)
        want, _ = orc.count_batch(qb, qo)
        assert np.array_equal(want, d_cnt[:sample].cpu().numpy().view(np.uint64))
        # host-buffer entry point on the same sample (pageable memory -> staged path)
        assert np.array_equal(ix.count_packed(qb, qo), want)
        del d_q, d_cnt
        # ---- locate at full size
        d_q, d_off = _device_queries(fxg, alphabet, n, text_seed, loc_nq, loc_len, 5)
        d_hoff = torch.zeros(loc_nq + 1, dtype=torch.int64, device="cuda")
        d_cnt = torch.zeros(loc_nq, dtype=torch.int64, device="cuda")
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), loc_nq, d_cnt.data_ptr(), st)
        assert ix.device_bytes()["full_sa"] == 4 * parts.bwt_len
        all_hits = []
        for variant in ((1, 2, 0) if alphabet == 0 else (1, 0)):   # LF-walk, bounded walk, unsampled-SA gather
            f.set_locate_variant(variant)
            try:
                ptr, n_hits = ix.locate_device(d_q.data_ptr(), d_off.data_ptr(), loc_nq, d_hoff.data_ptr(), stream=st)
            finally:
                f.set_locate_variant(0)
            assert n_hits == int(d_cnt.sum()) >= loc_nq         # locate total == count total
            hoff = d_hoff.cpu().numpy().view(np.uint64)
            assert np.array_equal(np.diff(hoff), d_cnt.cpu().numpy().view(np.uint64))
            buf = torch.empty(n_hits * 2, dtype=torch.int64, device="cuda")
            rc = C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), C.c_size_t(n_hits * 16), 3)
            assert rc == 0
            ix.device_free(ptr)
            all_hits.append(buf.cpu().numpy().view(np.uint64).reshape(-1, 2))
        hits = all_hits[-1]
        for other in all_hits[:-1]:
            assert np.array_equal(other, hits)              # all pass-2 variants agree on every hit
        assert int(hits[:, 0].max()) == 0 and int(hits[:, 1].max()) <= n - loc_len
        # round trip on a strided sample of hits: text at the hit == the query that produced it
        step = max(1, n_hits // 200_000)
        sel = np.arange(0, n_hits, step)
        qidx = np.searchsorted(hoff, sel, side="right") - 1
        windows = fx.gen_text_windows(alphabet, text_seed, hits[sel, 1], loc_len)
        queries = d_q.cpu().numpy().reshape(loc_nq, loc_len)[qidx]
        assert np.array_equal(windows, queries)
        # bit-exact vs the oracle on a prefix (same order: BWT rows)
        ns = min(sample, loc_nq)
        woff, whits, _ = orc.locate_batch(d_q[: ns * loc_len].cpu().numpy(),
                                          np.arange(ns + 1, dtype=np.uint64) * np.uint64(loc_len))
        assert np.array_equal(hoff[: ns + 1], woff) and np.array_equal(hits[: int(woff[-1])], whits)


def test_cfg2_cfg3_dna_3p1_gbp(fx, po):
    """cfg2: 10 M x 150 bp counts; cfg3: 1 M x 50 bp locate, SA ratio 8, k = 13, 3.1 Gbp text"""
    _run_config(fx, po, 0, 3_100_000_000, 13, 8, 10_000_000, 150, 1_000_000, 50, text_seed=3)


def test_cfg4_protein_2g_residues(fx, po):
    """cfg4: 10 M x 12-residue peptides, count + locate, k = 5, 2 G-residue text"""
    _run_config(fx, po, 1, 2_000_000_000, 5, 8, 10_000_000, 12, 1_000_000, 12, text_seed=6)


def test_cfg1_plumbing_1mbp_k13(fx, po):
    """cfg1 exactly: 10k random 32-bp queries on a 1 Mbp text, SA ratio 8, k = 13"""
    from conftest import device_from_parts, mixed_queries, oracle_from_parts
    text = fx.gen_text(0, 1_000_000, 1)
    parts = fx.build_parts(text, 0, ratio=8, kmer_len=13)
    orc = oracle_from_parts(po, parts)
    qb, qo = mixed_queries(fx, text, 10_000, 32, seed=2)
    with device_from_parts(parts) as ix:
        assert ix.kmer_len() == 13 and ix.device_bytes()["table"] == 8 * 4**13
        got = ix.count_packed(qb, qo)
        want, _ = orc.count_batch(qb, qo)
        assert np.array_equal(got, want) and int((want > 0).sum()) >= 5000
        off, hits = ix.locate_packed(qb, qo, sorted_hits=True)
        woff, whits, _ = orc.locate_batch(qb, qo, sorted_hits=True)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)


def test_cfg5_repeat_rich_1gbp_ratio32(po):
    """cfg5 at scale: 1 Gbp repeat-rich DNA, SA ratio 32, 1 M x 50-bp queries whose hit counts run from 1 to
    > 10^4.  The bounded walk (position-sampled array), the unsampled-SA gather and, on a prefix, the oracle
    agree on every hit; count totals equal hit totals."""
    import ctypes as C
    import torch
    from awry_b200 import FmIndex, fm_index as f
    from fixtures import pyfixture_gpu as fxg, repeats
    free, _ = torch.cuda.mem_get_info()
    if free < 100e9:
        pytest.skip("needs ~100 GB of free HBM")
    n, nq, L = 1_000_000_000, 1_000_000, 50
    fams = ((300, 100_010, 0.15), (6000, 20_005, 0.15))
    text, regions = repeats.repeat_rich_text(n, seed=8, tandem_arrays=40, families=fams)
    parts, _ = fxg.build_parts(0, n, 0, ratio=32, kmer_len=13, host_text=text)
    qb, qo = repeats.repeat_queries(text, regions, nq, L, seed=9)
    del text
    st = torch.cuda.current_stream().cuda_stream
    cudart = C.CDLL("libcudart.so")
    with FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words) as ix:
        assert ix.device_bytes()["lean_sa"] > 0 and ix.device_bytes()["full_sa"] > 0
        d_q = torch.from_numpy(qb).cuda()
        d_off = torch.from_numpy(qo.astype(np.int64)).cuda()
        d_cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
        d_hoff = torch.zeros(nq + 1, dtype=torch.int64, device="cuda")
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), nq, d_cnt.data_ptr(), st)
        total = int(d_cnt.sum())
        assert int(d_cnt.max()) > 10_000 and int(d_cnt.min()) == 1
        bufs = []
        for variant in (2, 0):
            f.set_locate_variant(variant)
            try:
                ptr, n_hits = ix.locate_device(d_q.data_ptr(), d_off.data_ptr(), nq, d_hoff.data_ptr(), stream=st)
            finally:
                f.set_locate_variant(0)
            assert n_hits == total
            assert torch.equal(d_hoff[1:] - d_hoff[:-1], d_cnt)
            buf = torch.empty(n_hits * 2, dtype=torch.int64, device="cuda")
            assert cudart.cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(ptr), C.c_size_t(n_hits * 16), 3) == 0
            ix.device_free(ptr)
            bufs.append(buf)
        assert torch.equal(bufs[0], bufs[1])
        ns = 2000
        orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                        parts.prefix_sums, parts.sa_words)
        woff, whits, _ = orc.locate_batch(qb[: ns * L], qo[: ns + 1])
        assert np.array_equal(d_hoff[: ns + 1].cpu().numpy().view(np.uint64), woff)
        got = bufs[0][: 2 * int(woff[-1])].cpu().numpy().view(np.uint64).reshape(-1, 2)
        assert np.array_equal(got, whits)
