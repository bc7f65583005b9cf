"""The GPU fixture builder (bench-scale index construction) against the CPU builder."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alphabet,n,ratio", [(0, 1_000_003, 8), (0, 70_001, 5), (1, 300_000, 8), (1, 999, 3)])
def test_gpu_builder_matches_cpu_builder(fx, alphabet, n, ratio):
    from fixtures import pyfixture_gpu as fxg
    text = fx.gen_text(alphabet, n, 77)
    want = fx.build_parts(text, alphabet, ratio=ratio, kmer_len=4)
    got, phases = fxg.build_parts(alphabet, n, 77, ratio=ratio, kmer_len=4)
    assert np.array_equal(got.prefix_sums, want.prefix_sums)
    assert np.array_equal(got.sa_words, want.sa_words)
    assert np.array_equal(got.blocks, want.blocks)
    # same, with the text uploaded from the host instead of regenerated from the seed
    got2, _ = fxg.build_parts(alphabet, n, 0, ratio=ratio, kmer_len=4, host_text=text)
    assert np.array_equal(got2.blocks, want.blocks) and np.array_equal(got2.sa_words, want.sa_words)


def test_gpu_builder_with_ties_and_ambiguity(fx):
    """poly-A runs and N's: 32-symbol keys tie, the host fix-up must order them like the CPU sorter"""
    from fixtures import pyfixture_gpu as fxg
    text = bytearray(fx.gen_text(0, 50_000, 5).tobytes())
    text[1000:1100] = b"A" * 100
    text[30000:30090] = b"A" * 90
    text[-40:] = b"A" * 40
    t1 = np.frombuffer(bytes(text), dtype=np.uint8)
    want = fx.build_parts(t1, 0, ratio=4, kmer_len=4)
    got, _ = fxg.build_parts(0, len(t1), 0, ratio=4, kmer_len=4, host_text=t1)
    assert np.array_equal(got.blocks, want.blocks) and np.array_equal(got.sa_words, want.sa_words)
    text[2000:2010] = b"NNNNNNNNNN"
    text[777] = ord("N")
    t2 = np.frombuffer(bytes(text), dtype=np.uint8)
    want = fx.build_parts(t2, 0, ratio=4, kmer_len=4)
    got, _ = fxg.build_parts(0, len(t2), 0, ratio=4, kmer_len=4, host_text=t2)
    assert np.array_equal(got.blocks, want.blocks) and np.array_equal(got.sa_words, want.sa_words)
    assert np.array_equal(got.prefix_sums, want.prefix_sums)


def test_device_query_generator_matches_cpu(fx):
    import torch
    from fixtures import pyfixture_gpu as fxg
    n, nq, qlen = 100_000, 5000, 37
    text = fx.gen_text(0, n, 9)
    qb, _, pos = fx.gen_substring_queries(text, nq, qlen, 4)
    d = torch.empty(nq * qlen, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(0, n, 9, nq, qlen, 4, d.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), qb)
    fxg.gen_queries_device(0, n, 9, nq, qlen, 4, d.data_ptr(), mut_ppm=100_000)
    torch.cuda.synchronize()
    diff = (d.cpu().numpy() != qb).reshape(nq, qlen).sum(axis=1)
    assert diff.max() == 1 and 300 < int(diff.sum()) < 700


@pytest.mark.parametrize("alphabet", [0, 1])
def test_gpu_prefix_doubling_matches_cpu_builder(fx, alphabet, monkeypatch):
    """repeat-rich text through the GPU prefix-doubling path (forced), against the CPU sorter"""
    from fixtures import pyfixture_gpu as fxg, repeats
    monkeypatch.setenv("AWRY_FIXTURE_HOST_FIXUP_MAX", "0")
    if alphabet == 0:
        text, _ = repeats.repeat_rich_text(1_500_000, seed=3)
        text[1000:1040] = ord("N")
    else:
        rng = np.random.default_rng(4)
        aa = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
        text = aa[rng.integers(0, 20, 400_000)]
        unit = aa[rng.integers(0, 20, 37)]
        text[5000:5000 + 37 * 300] = np.tile(unit, 300)
        text[200_000:200_000 + 37 * 100] = np.tile(unit, 100)
    text = np.ascontiguousarray(text)
    want = fx.build_parts(text, alphabet, ratio=8, kmer_len=4)
    got, phases = fxg.build_parts(alphabet, len(text), 0, ratio=8, kmer_len=4, host_text=text)
    assert np.array_equal(got.prefix_sums, want.prefix_sums)
    assert np.array_equal(got.sa_words, want.sa_words)
    assert np.array_equal(got.blocks, want.blocks)
