"""GPU index construction (awry_build_parts / awry_index_build / awry_build_index_file, SURVEY.md 8(f)
rank 2) against the independent CPU builder and file writer of fixtures/ (the restatement of
FmIndex::new + save, fm_index.rs:202-268, fm_index_file.rs:42-106)."""
import os

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
    monkeypatch.setenv("AWRY_B200_BUILD_HOST_FIXUP_MAX", "0")
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


def _write_fasta(path, records, headers, width=60, lower_every=3, crlf=False):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for i, (h, r) in enumerate(zip(headers, records)):
            f.write(">" + h + nl)
            r = r.lower() if lower_every and i % lower_every == 1 else r
            for j in range(0, len(r), width):
                f.write(r[j:j + width] + nl)


def _write_fastq(path, records, headers):
    with open(path, "w") as f:
        for h, r in zip(headers, records):
            f.write(f"@{h}\n{r}\n+\n{'I' * len(r)}\n")


def _dna_records(fx, lens, seed):
    recs = []
    for i, n in enumerate(lens):
        t = bytearray(fx.gen_text(0, n, seed + i).tobytes())
        if n > 200:
            t[50:53] = b"NNN"          # ambiguity inside a record
        recs.append(t.decode())
    return recs


@pytest.mark.parametrize("kind", ["fasta", "fasta_crlf", "fastq"])
def test_build_index_file_is_byte_identical_to_cpu_writer(fx, po, tmp_path, kind):
    """FASTA/FASTQ -> .awry on the GPU == the CPU restatement of FmIndex::new + save, byte for byte
    (blocks, prefix sums, SA words, the reference-style k-mer table section, sequence index)."""
    from awry_b200 import FmBuildArgs, FmIndex, fm_index as f
    recs = _dna_records(fx, [5000, 1, 777, 12_345, 64], 11)
    headers = ["chr1 test record", "tiny", "chr3", "chr4 len=12345", "last"]
    src = str(tmp_path / ("in.fq" if kind == "fastq" else "in.fa"))
    if kind == "fastq":
        _write_fastq(src, recs, headers)
    else:
        _write_fasta(src, recs, headers, crlf=kind == "fasta_crlf")
    text, starts = fx.concat_records(recs, 0)
    want = fx.build_parts(text, 0, ratio=5, kmer_len=6, seq_starts=starts, headers=headers)
    want_path = want.write(str(tmp_path / "want.awry"))
    got_path = str(tmp_path / "got.awry")
    f.build_index_file(src, got_path, 0, suffix_array_compression_ratio=5, lookup_table_kmer_len=6)
    assert open(got_path, "rb").read() == open(want_path, "rb").read()
    # FmIndex::new straight to a searchable index (+ save in the same call), then search it
    got2 = str(tmp_path / "got2.awry")
    args = FmBuildArgs(src, suffix_array_compression_ratio=5, lookup_table_kmer_len=6, alphabet=0)
    orc = po.OracleIndex.load(want_path)
    qs = [text[i:i + 20].tobytes() for i in range(0, len(text) - 20, 37)] + [b"ACGTACGTAC", b"NNN", b"N"]
    with FmIndex.new(args, save_to=got2) as ix, FmIndex.load(got_path) as ix2:
        assert open(got2, "rb").read() == open(want_path, "rb").read()
        assert ix.bwt_len() == len(text) + 1 and ix.kmer_len() == 6
        assert [ix.sequence_header(i) for i in range(5)] == headers
        qb, qo = f.pack_queries(qs)
        wc, _ = orc.count_batch(qb, qo)
        assert np.array_equal(ix.count_packed(qb, qo), wc) and np.array_equal(ix2.count_packed(qb, qo), wc)
        woff, whits, _ = orc.locate_batch(qb, qo)
        off, hits = ix.locate_packed(qb, qo)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    os.remove(got2)


def test_build_index_file_amino_defaults(fx, po, tmp_path):
    from awry_b200 import fm_index as f
    recs = [fx.gen_text(1, n, 21 + i).tobytes().decode() for i, n in enumerate([3000, 500])]
    recs[1] = recs[1][:100] + "BZJ" + recs[1][103:]          # letters outside the alphabet -> X (alphabet.rs:199-222)
    headers = ["sp|P1", "sp|P2"]
    src = str(tmp_path / "p.fa")
    _write_fasta(src, recs, headers, width=70, lower_every=0)
    text, starts = fx.concat_records(recs, 1)
    want = fx.build_parts(text, 1, seq_starts=starts, headers=headers)      # defaults: ratio 8, k 4
    want_path = want.write(str(tmp_path / "want.awry"))
    got_path = f.build_index_file(src, str(tmp_path / "got.awry"), 1)
    assert open(got_path, "rb").read() == open(want_path, "rb").read()


def test_build_errors(tmp_path):
    from awry_b200 import AwryError, fm_index as f
    with pytest.raises(AwryError) as e:
        f.build_index_file(str(tmp_path / "missing.fa"), str(tmp_path / "o.awry"))
    assert e.value.code == -2
    bad = tmp_path / "bad.txt"
    bad.write_text("ACGT\nACGT\n")
    with pytest.raises(AwryError) as e:
        f.build_index_file(str(bad), str(tmp_path / "o.awry"))
    assert e.value.code == -2 and "FASTA" in str(e.value)
    dollar = tmp_path / "dollar.fa"
    dollar.write_text(">x\nACGT$ACGT\n")
    with pytest.raises(AwryError):
        f.build_index_file(str(dollar), str(tmp_path / "o.awry"))
    empty = tmp_path / "empty.fa"
    empty.write_text(">only a header\n")
    with pytest.raises(AwryError):
        f.build_index_file(str(empty), str(tmp_path / "o.awry"))


@pytest.mark.parametrize("seed", range(12))
def test_random_sequence_files_build_byte_identical(fx, tmp_path, monkeypatch, seed):
    """random multi-record FASTA files (both alphabets, ambiguity letters, repeats, tiny and empty-ish records,
    host fix-up and forced GPU prefix doubling, multi-threaded FASTA parse) -> same `.awry` bytes as the CPU
    restatement of FmIndex::new + save"""
    from awry_b200 import fm_index as f
    rng = np.random.default_rng(500 + seed)
    alphabet = seed % 2
    letters = b"ACGT" if alphabet == 0 else b"ACDEFGHIKLMNPQRSTVWY"
    amb = b"NRYKM" if alphabet == 0 else b"XBZJ"
    recs = []
    for _ in range(int(rng.integers(1, 7))):
        ln = int(rng.choice([1, 2, 5, 64, 255, 256, 257, 3000, 20000]))
        r = bytearray(np.frombuffer(letters, dtype=np.uint8)[rng.integers(0, len(letters), ln)].tobytes())
        if ln > 500 and rng.random() < 0.6:                      # tandem repeat: long ties
            unit = bytes(r[:int(rng.integers(1, 40))])
            reps = int(rng.integers(5, 60))
            r[100:100 + len(unit) * reps] = (unit * reps)[: max(0, min(len(r) - 100, len(unit) * reps))]
        for _ in range(int(rng.integers(0, 4))):
            r[int(rng.integers(0, len(r)))] = amb[int(rng.integers(0, len(amb)))]
        recs.append(bytes(r).decode())
    headers = [f"r{i} seed {seed}" for i in range(len(recs))]
    src = str(tmp_path / "in.fa")
    _write_fasta(src, recs, headers, width=int(rng.integers(1, 100)), lower_every=2 if seed % 3 == 0 else 0)
    ratio, k = int(rng.choice([1, 3, 8, 32])), int(rng.integers(1, 7 if alphabet == 0 else 4))
    text, starts = fx.concat_records(recs, alphabet)
    want = fx.build_parts(text, alphabet, ratio=ratio, kmer_len=k, seq_starts=starts, headers=headers)
    want_path = want.write(str(tmp_path / "want.awry"))
    if seed % 2 == 0:
        monkeypatch.setenv("AWRY_B200_BUILD_HOST_FIXUP_MAX", "0")      # GPU prefix doubling for every tie
    monkeypatch.setenv("AWRY_B200_FASTA_RANGE_BYTES", "300")           # several parser threads on a small file
    got_path = f.build_index_file(src, str(tmp_path / "got.awry"), alphabet, suffix_array_compression_ratio=ratio,
                                  lookup_table_kmer_len=k)
    assert open(got_path, "rb").read() == open(want_path, "rb").read()
