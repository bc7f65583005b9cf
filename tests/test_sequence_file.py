"""The builder's FASTA / FASTQ reader (awry_read_sequence_file; host only) against the Python model of
libsufr's read_sequence_file that the fixtures use (records joined by 'N' / 'X', fm_index.rs:148-153),
with the multi-threaded FASTA path forced onto small files."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from awry_b200 import fm_index as f
from fixtures import pyfixture as fx
rng = np.random.default_rng(3)
tmp = sys.argv[1]
def fasta(path, recs, hdrs, width, crlf, lead_blank, last_nl):
    nl = b"\r\n" if crlf else b"\n"
    with open(path, "wb") as fh:
        if lead_blank: fh.write(nl + nl)
        for i, (h, r) in enumerate(zip(hdrs, recs)):
            fh.write(b">" + h + nl)
            for j in range(0, len(r), width):
                last = i + 1 == len(recs) and j + width >= len(r)
                fh.write(r[j:j + width] + (nl if (last_nl or not last) else b""))
for alphabet, letters in ((0, b"ACGTacgtN"), (1, b"ACDEFGHIKLMNPQRSTVWY")):
    for trial in range(12):
        nrec = int(rng.integers(1, 9))
        recs = [bytes(np.frombuffer(letters, dtype=np.uint8)[rng.integers(0, len(letters), int(rng.integers(0 if nrec > 1 else 1, 3000)))]) for _ in range(nrec)]
        if sum(len(r) for r in recs) == 0: recs[0] = b"ACGT"
        hdrs = [b"rec%%d some text" %% i for i in range(nrec)]
        path = tmp + "/t.fa"
        fasta(path, recs, hdrs, int(rng.integers(1, 120)), trial %% 3 == 0, trial %% 4 == 0, trial %% 5 != 0)
        want, wstarts = fx.concat_records([r.upper() for r in recs], alphabet)
        text, starts = f.read_sequence_file(path, alphabet)
        assert np.array_equal(np.frombuffer(text.tobytes().upper(), dtype=np.uint8), want), (alphabet, trial)
        assert np.array_equal(starts, wstarts), (alphabet, trial, starts, wstarts)
    # FASTQ
    recs = [bytes(np.frombuffer(letters, dtype=np.uint8)[rng.integers(0, len(letters), int(rng.integers(1, 400)))]) for _ in range(50)]
    path = tmp + "/t.fq"
    with open(path, "wb") as fh:
        for i, r in enumerate(recs):
            fh.write(b"@r%%d\n" %% i + r + b"\n+\n" + b"@" * len(r) + b"\n")
    want, wstarts = fx.concat_records([r.upper() for r in recs], alphabet)
    text, starts = f.read_sequence_file(path, alphabet)
    assert np.array_equal(np.frombuffer(text.tobytes().upper(), dtype=np.uint8), want) and np.array_equal(starts, wstarts)
    # gzip-compressed input goes through the same parser
    import gzip
    with open(path, "rb") as fi, gzip.open(path + ".gz", "wb", compresslevel=1) as fo:
        fo.write(fi.read())
    t2, s2 = f.read_sequence_file(path + ".gz", alphabet)
    assert np.array_equal(t2, text) and np.array_equal(s2, starts)
print("ok")
""" % ROOT


@pytest.mark.parametrize("range_bytes", ["16", "97", "1000", str(32 << 20)])
def test_reader_matches_model(tmp_path, range_bytes):
    env = dict(os.environ, AWRY_B200_FASTA_RANGE_BYTES=range_bytes)
    out = subprocess.run([sys.executable, "-c", CHECK, str(tmp_path)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-3000:]


def test_reader_errors(tmp_path):
    from awry_b200 import AwryError, fm_index as f
    with pytest.raises(AwryError) as e:
        f.read_sequence_file(str(tmp_path / "missing.fa"))
    assert e.value.code == -2
    p = tmp_path / "x.txt"
    p.write_text("ACGT\n")
    with pytest.raises(AwryError) as e:
        f.read_sequence_file(str(p))
    assert e.value.code == -2 and "FASTA" in str(e.value)
    p.write_text("\n\n")
    with pytest.raises(AwryError):
        f.read_sequence_file(str(p))
    p.write_text(">h\n")
    with pytest.raises(AwryError):
        f.read_sequence_file(str(p))
