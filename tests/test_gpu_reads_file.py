"""Streaming reads-file front-end (awry_count_reads_file / awry_locate_reads_file):
FASTQ / FASTA parsed on the device must give exactly what parallel_count / parallel_locate give on the
same reads parsed by Python, for every chunk size (records straddling chunk boundaries are carried)."""
import numpy as np
import pytest

from conftest import device_from_parts, oracle_from_parts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dna(fx):
    text = fx.gen_text(0, 150_000, 31)
    return fx.build_parts(text, 0, ratio=8, kmer_len=7)


@pytest.fixture(scope="module")
def dev(dna):
    ix = device_from_parts(dna)
    yield ix
    ix.close()


def _reads(fx, dna, n, seed, ragged=True):
    rng = np.random.default_rng(seed)
    text = bytes(dna.text)
    out = []
    for i in range(n):
        ln = int(rng.integers(1, 260)) if ragged else 100
        p = int(rng.integers(0, len(text) - ln))
        r = bytearray(text[p:p + ln])
        k = i % 7
        if k == 3:
            r[len(r) // 2] = ord("N")
        elif k == 5:
            r = bytearray(bytes(r).lower())
        elif k == 6 and ln > 10:
            r[3] = ord("ACGT"[(r[3] % 4)])          # likely a mismatch
        out.append(bytes(r))
    return out


def _write_fastq(path, reads, crlf=False, final_newline=True):
    nl = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            q = bytes((33 + (i * 7 + j) % 41) for j in range(len(r)))     # qualities include '@' and '>'
            q = b"@" + q[1:] if i % 5 == 0 and len(q) > 1 else q
            rec = b"@read" + str(i).encode() + b" some description" + nl + r + nl + b"+" + nl + q
            f.write(rec + (nl if (final_newline or i + 1 < len(reads)) else b""))


def _write_fasta(path, reads, width=70, crlf=False):
    nl = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as f:
        f.write(nl)                                      # leading blank line
        for i, r in enumerate(reads):
            f.write(b">r" + str(i).encode() + nl)
            w = width if i % 3 else len(r) + 1           # every third record on one line
            for j in range(0, len(r), w):
                f.write(r[j:j + w] + nl)


def _expect(po, dna, dev, reads):
    from awry_b200 import fm_index as f
    orc = oracle_from_parts(po, dna)
    qb, qo = f.pack_queries(reads)
    wc, _ = orc.count_batch(qb, qo)
    woff, whits, _ = orc.locate_batch(qb, qo)
    return wc, woff, whits


@pytest.mark.parametrize("chunk", [700, 4096, 1 << 16, 64 << 20])
@pytest.mark.parametrize("kind", ["fastq", "fastq_crlf_nonl", "fasta", "fasta_crlf"])
def test_reads_file_matches_batch_api(fx, po, dna, dev, tmp_path, monkeypatch, chunk, kind):
    reads = _reads(fx, dna, 3000, seed=len(kind) + chunk % 97)
    path = str(tmp_path / "reads.txt")
    if kind.startswith("fastq"):
        _write_fastq(path, reads, crlf="crlf" in kind, final_newline="nonl" not in kind)
    else:
        _write_fasta(path, reads, crlf="crlf" in kind)
    wc, woff, whits = _expect(po, dna, dev, reads)
    monkeypatch.setenv("AWRY_B200_READS_CHUNK", str(chunk))
    got = dev.count_reads_file(path)
    assert got.shape == wc.shape and np.array_equal(got, wc)
    off, hits = dev.locate_reads_file(path)
    assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    if chunk == 4096:
        soff, shits = dev.locate_reads_file(path, sorted_hits=True)
        orc = oracle_from_parts(po, dna)
        from awry_b200 import fm_index as f
        qb, qo = f.pack_queries(reads)
        woff2, whits2, _ = orc.locate_batch(qb, qo, sorted_hits=True)
        assert np.array_equal(soff, woff2) and np.array_equal(shits, whits2)


def test_reads_file_larger_than_one_chunk_default_settings(fx, po, dna, dev, tmp_path):
    """~100 MB FASTQ through the default 64 MiB chunks (two chunks + a carried record)"""
    reads = _reads(fx, dna, 400_000, seed=5, ragged=False)
    path = str(tmp_path / "big.fq")
    _write_fastq(path, reads)
    from awry_b200 import fm_index as f
    qb, qo = f.pack_queries(reads)
    want = dev.count_packed(qb, qo)
    got = dev.count_reads_file(path)
    assert np.array_equal(got, want)
    off, hits = dev.locate_reads_file(path)
    woff, whits = dev.locate_packed(qb, qo)
    assert np.array_equal(off, woff) and np.array_equal(hits, whits)


def test_reads_file_errors(dev, tmp_path, monkeypatch):
    from awry_b200 import AwryError
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(tmp_path / "missing.fq"))
    assert e.value.code == -2
    p = tmp_path / "x.txt"
    p.write_bytes(b"ACGT\nACGT\n")
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(p))
    assert e.value.code == -3
    p.write_bytes(b"\x1f\x8b\x08\x00garbage")                      # gzip magic, corrupt stream
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(p))
    assert e.value.code == -3
    p.write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\nACGT\n+\n")           # truncated last record
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(p))
    assert e.value.code == -3 and "truncated" in str(e.value)
    p.write_bytes(b"@r1\nACGT\n+\nIIII\n@r2\n\n+\n\n")               # empty read: the reference would panic
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(p))
    assert e.value.code == -5 and "read 1" in str(e.value)
    p.write_bytes(b"")
    assert len(dev.count_reads_file(str(p))) == 0
    off, hits = dev.locate_reads_file(str(p))
    assert list(off) == [0] and len(hits) == 0
    p.write_bytes(b"@r1\n" + b"A" * 5000 + b"\n+\n" + b"I" * 5000 + b"\n")
    monkeypatch.setenv("AWRY_B200_READS_CHUNK", "1024")               # a record larger than the carry buffer
    with pytest.raises(AwryError) as e:
        dev.count_reads_file(str(p))
    assert e.value.code == -6
    p.write_bytes(b"@r1\nACGT\n+\nIIII\n\n\n")                         # trailing blank lines are fine
    assert list(dev.count_reads_file(str(p))) == [int(dev.count_string("ACGT"))]


@pytest.mark.parametrize("chunk", [900, 1 << 16, 64 << 20])
def test_gzip_compressed_reads(fx, po, dna, dev, tmp_path, monkeypatch, chunk):
    """reads.fastq.gz / reads.fa.gz: inflated by the reader thread, everything after that unchanged"""
    import gzip
    reads = _reads(fx, dna, 4000, seed=77)
    wc, woff, whits = _expect(po, dna, dev, reads)
    monkeypatch.setenv("AWRY_B200_READS_CHUNK", str(chunk))
    for kind in ("fastq", "fasta"):
        plain = str(tmp_path / f"r.{kind}")
        (_write_fastq if kind == "fastq" else _write_fasta)(plain, reads)
        gzp = plain + ".gz"
        with open(plain, "rb") as fi, gzip.open(gzp, "wb", compresslevel=1) as fo:
            fo.write(fi.read())
        assert np.array_equal(dev.count_reads_file(gzp), wc)
        off, hits = dev.locate_reads_file(gzp)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
    empty = str(tmp_path / "empty.fq.gz")
    with gzip.open(empty, "wb") as fo:
        fo.write(b"")
    assert len(dev.count_reads_file(empty)) == 0


@pytest.mark.parametrize("kind", ["fastq", "fasta"])
def test_reads_file_over_several_replicas(tmp_path, fx, po, dna, monkeypatch, kind):
    """a plain reads file on a handle with several replicas: the file is cut at record starts found on the host
    (quality lines that begin with '@' must not fool it), every replica parses and searches its own segment,
    the results come back in file order and equal the single-replica run and the oracle"""
    import torch
    from awry_b200 import AwryError, fm_index as f
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs >= 2 GPUs")
    reads = _reads(fx, dna, 30_000, seed=5)
    path = str(tmp_path / ("r." + kind))
    if kind == "fastq":
        _write_fastq(path, reads)
    else:
        _write_fasta(path, reads)
    orc = oracle_from_parts(po, dna)
    qb, qo = f.pack_queries(reads)
    want, _ = orc.count_batch(qb, qo)
    woff, whits, _ = orc.locate_batch(qb, qo)
    monkeypatch.setenv("AWRY_B200_READS_MIN_SEGMENT", "200000")       # (default 128 MiB: this file is a few MB)
    with device_from_parts(dna, devices=list(range(min(n_dev, 8)))) as ix:
        got = ix.count_reads_file(path)
        assert np.array_equal(got, want)
        off, hits = ix.locate_reads_file(path)
        assert np.array_equal(off, woff) and np.array_equal(hits, whits)
        monkeypatch.setenv("AWRY_B200_READS_REPLICAS", "1")            # the same handle, replica 0 only
        assert np.array_equal(ix.count_reads_file(path), want)
        monkeypatch.delenv("AWRY_B200_READS_REPLICAS")
        # a sentinel in a late read is reported with its file-wide read number
        bad = list(reads)
        bad[27_123] = bad[27_123][:5] + b"$" + bad[27_123][6:] if len(bad[27_123]) > 6 else b"AC$GT"
        bpath = str(tmp_path / ("bad." + kind))
        (_write_fastq if kind == "fastq" else _write_fasta)(bpath, bad)
        with pytest.raises(AwryError) as e:
            ix.count_reads_file(bpath)
        assert e.value.code == -5 and "read 27123 " in str(e.value)
