"""FmIndex::save (fm_index_file.rs:42-106) of a handle that only holds the device layout
(awry_index_save): the bwt.rs blocks, milestones and the reference-style k-mer table section are re-derived
on the device and must reproduce, byte for byte, the file the CPU restatement of FmIndex::new + save writes
(fixtures/) -- the reference's own save -> load equality test (fm_index.rs:1046-1088) taken one step
further: load -> save -> identical file."""
import hashlib
import os

import numpy as np
import pytest

from conftest import device_from_parts

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "appendix_a.awry")


def _records(fx, alphabet, lens, seed):
    recs = []
    for i, n in enumerate(lens):
        t = bytearray(fx.gen_text(alphabet, n, seed + i).tobytes())
        if n > 200:
            t[50:53] = b"NNN" if alphabet == 0 else b"XXX"      # ambiguity symbols inside a record
        recs.append(t.decode())
    return recs


@pytest.mark.parametrize("alphabet,lens,ratio,k", [
    (0, [5000, 1, 777, 12_345, 64], 5, 6),       # bwt_len not a multiple of 256 / 128: padded last block
    (0, [255], 1, 1),                            # exactly one 256-row reference block (255 symbols + '$')
    (0, [256], 8, 3),                            # one symbol into the second block
    (0, [300_000, 70_001], 8, 10),               # several hundred blocks, the reference's default k
    (1, [3000, 500], 8, 4),
    (1, [63], 3, 2),
    (1, [40_000, 9, 130], 32, 3),
])
def test_load_then_save_reproduces_the_file(fx, tmp_path, alphabet, lens, ratio, k):
    from awry_b200 import FmIndex
    recs = _records(fx, alphabet, lens, 31)
    headers = [f"rec{i} len={n}" for i, n in enumerate(lens)]
    headers[-1] = ""                                             # an empty header is legal (header_len 0)
    text, starts = fx.concat_records(recs, alphabet)
    want = fx.build_parts(text, alphabet, ratio=ratio, kmer_len=k, seq_starts=starts, headers=headers)
    want_path = want.write(str(tmp_path / "want.awry"))
    want_bytes = open(want_path, "rb").read()
    # a handle loaded from the file
    got_path = str(tmp_path / "got.awry")
    with FmIndex.load(want_path) as ix:
        ix.save(got_path)
        assert open(got_path, "rb").read() == want_bytes
        ix.save(got_path)                                        # an existing file is truncated and rewritten
        assert open(got_path, "rb").read() == want_bytes
    # a handle made from the arrays FmIndex::new hands over
    got2 = str(tmp_path / "got2.awry")
    with device_from_parts(want) as ix:
        ix.save(got2)
    assert open(got2, "rb").read() == want_bytes
    # and the saved file loads and answers like the original
    with FmIndex.load(got_path) as a, FmIndex.load(want_path) as b:
        qs = [text[i:i + 12].tobytes() for i in range(0, max(1, len(text) - 12), 97)]
        assert np.array_equal(a.parallel_count(qs), b.parallel_count(qs))
        assert a.kmer_len() == k and a.suffix_array_compression_ratio() == ratio
        assert [a.sequence_header(i) for i in range(len(lens))] == headers


def test_save_of_a_gpu_built_index(fx, tmp_path):
    """FmIndex::new on the GPU, then save as a separate call == new + save in one call == the CPU writer"""
    from awry_b200 import FmBuildArgs, FmIndex
    recs = _records(fx, 0, [4000, 333], 7)
    src = str(tmp_path / "in.fa")
    with open(src, "w") as f:
        for i, r in enumerate(recs):
            f.write(f">r{i}\n{r}\n")
    text, starts = fx.concat_records(recs, 0)
    want = fx.build_parts(text, 0, ratio=4, kmer_len=5, seq_starts=starts, headers=["r0", "r1"])
    want_bytes = open(want.write(str(tmp_path / "want.awry")), "rb").read()
    args = FmBuildArgs(src, suffix_array_compression_ratio=4, lookup_table_kmer_len=5, alphabet=0)
    with FmIndex.new(args) as ix:
        ix.save(str(tmp_path / "got.awry"))
    assert open(tmp_path / "got.awry", "rb").read() == want_bytes


def test_save_golden_appendix_a(tmp_path):
    """the 557-byte worked example with its pinned sha256"""
    from awry_b200 import FmIndex
    out = str(tmp_path / "a.awry")
    with FmIndex.load(GOLDEN) as ix:
        ix.save(out)
    data = open(out, "rb").read()
    assert len(data) == 557
    assert hashlib.sha256(data).hexdigest() == "bb57bc33cfedac88a1ccbd980b8af3263be80d1bbb1d4954fb836aa0c26c5f3e"


def test_save_errors(fx, tmp_path):
    from awry_b200 import AwryError, FmIndex
    text = fx.gen_text(0, 2000, 3)
    parts = fx.build_parts(text, 0, ratio=8, kmer_len=3)
    with device_from_parts(parts) as ix:
        with pytest.raises(AwryError) as e:
            ix.save(str(tmp_path / "no_such_dir" / "x.awry"))
        assert e.value.code == -2
        ix.save(str(tmp_path / "ok.awry"))                      # the handle is still usable afterwards
        assert np.array_equal(ix.parallel_count([text[:20].tobytes()]), [1])


@pytest.mark.parametrize("alphabet,n", [(0, 70_000_123), (1, 68_000_001)])
def test_save_streams_more_than_one_chunk_of_blocks(fx, tmp_path, alphabet, n):
    """more than 2^18 reference blocks: the inverse re-layout runs chunk by chunk through the device and pinned
    double buffers (first block of a chunk > 0, buffers reused) -- still the CPU writer's file, byte for byte"""
    from awry_b200 import FmIndex
    from fixtures import pyfixture_gpu as fxg
    parts, _ = fxg.build_parts(alphabet, n, 5, ratio=16, kmer_len=3)     # index arrays by the GPU builder
    assert len(parts.blocks) // (20 if alphabet == 0 else 44) > (1 << 18)
    want_path = parts.write(str(tmp_path / "want.awry"))                # file by the CPU writer
    got_path = str(tmp_path / "got.awry")
    with FmIndex.load(want_path) as ix:
        ix.save(got_path)
    a, b = np.fromfile(want_path, dtype=np.uint8), np.fromfile(got_path, dtype=np.uint8)
    assert a.shape == b.shape and np.array_equal(a, b)
