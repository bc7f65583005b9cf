"""ctypes binding of the CPU fixture builder (fixtures/libawry_fixture_cpu.so).

Builds reference-format index parts / `.awry` v1 files from a text, standing in for the
reference's FmIndex::new + save (which need a Rust toolchain and libsufr).  Test/bench
infrastructure only -- the product package never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

NUCLEOTIDE, AMINO = 0, 1


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libawry_fixture_cpu.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, u64 = C.c_void_p, C.c_uint64
        L.fx_last_error.restype = C.c_char_p
        L.fx_gen_text.argtypes = [C.c_int, u64, u64, vp]
        L.fx_gen_text.restype = None
        L.fx_gen_substring_queries.argtypes = [vp, u64, u64, u64, u64, vp, vp]
        L.fx_gen_substring_queries.restype = None
        L.fx_gen_text_windows.argtypes = [C.c_int, u64, vp, u64, u64, vp]
        L.fx_gen_text_windows.restype = None
        for name in ("fx_num_blocks",):
            getattr(L, name).argtypes = [u64]
            getattr(L, name).restype = u64
        L.fx_block_words.argtypes = [C.c_int]
        L.fx_block_words.restype = u64
        L.fx_sa_words.argtypes = [u64, u64]
        L.fx_sa_words.restype = u64
        L.fx_table_entries.argtypes = [C.c_int, C.c_uint]
        L.fx_table_entries.restype = u64
        L.fx_suffix_array.argtypes = [C.c_int, vp, u64, vp]
        L.fx_build_parts.argtypes = [C.c_int, vp, u64, vp, u64, vp, vp, vp]
        L.fx_populate_kmer_table.argtypes = [C.c_int, u64, vp, vp, C.c_uint, vp]
        L.fx_populate_kmer_table.restype = None
        L.fx_write_awry.argtypes = [C.c_char_p, C.c_int, u64, u64, vp, vp, vp, C.c_uint, vp, vp,
                                    vp, u64]
        _LIB = L
    return _LIB


def gen_text(alphabet, n, seed):
    out = np.empty(n, dtype=np.uint8)
    lib().fx_gen_text(alphabet, n, seed, out.ctypes.data)
    return out


def gen_substring_queries(text, nq, qlen, seed):
    """-> (qbytes uint8[nq*qlen], qoff uint64[nq+1], positions uint64[nq])"""
    text = np.ascontiguousarray(text, dtype=np.uint8)
    qbytes = np.empty(nq * qlen, dtype=np.uint8)
    pos = np.empty(nq, dtype=np.uint64)
    lib().fx_gen_substring_queries(text.ctypes.data, len(text), nq, qlen, seed, qbytes.ctypes.data,
                                   pos.ctypes.data)
    qoff = np.arange(nq + 1, dtype=np.uint64) * np.uint64(qlen)
    return qbytes, qoff, pos


def gen_text_windows(alphabet, seed, positions, length):
    """text[p : p+length] of the synthetic text for every p in positions -> uint8[len(positions), length]"""
    positions = np.ascontiguousarray(positions, dtype=np.uint64)
    out = np.empty((len(positions), length), dtype=np.uint8)
    lib().fx_gen_text_windows(alphabet, seed, positions.ctypes.data, len(positions), length, out.ctypes.data)
    return out


class Parts:
    """Reference-layout index arrays (what FmIndex::new produces, fm_index.rs:242-251)."""

    def __init__(self, alphabet, ratio, bwt_len, kmer_len, blocks, prefix_sums, sa_words,
                 seq_starts, headers, text, sa=None):
        self.alphabet, self.ratio, self.bwt_len, self.kmer_len = alphabet, ratio, bwt_len, kmer_len
        self.blocks, self.prefix_sums, self.sa_words = blocks, prefix_sums, sa_words
        self.seq_starts, self.headers, self.text, self.sa = seq_starts, headers, text, sa

    def write(self, path, reference_table=True):
        L = lib()
        table_ptr = None
        table = None
        if reference_table:
            n_entries = L.fx_table_entries(self.alphabet, self.kmer_len)
            table = np.empty(2 * n_entries, dtype=np.uint64)
            L.fx_populate_kmer_table(self.alphabet, self.bwt_len, self.blocks.ctypes.data,
                                     self.prefix_sums.ctypes.data, self.kmer_len, table.ctypes.data)
            table_ptr = table.ctypes.data
        hdrs = (C.c_char_p * len(self.headers))(*[h.encode() for h in self.headers])
        rc = L.fx_write_awry(os.fsencode(path), self.alphabet, self.ratio, self.bwt_len,
                             self.blocks.ctypes.data, self.prefix_sums.ctypes.data,
                             self.sa_words.ctypes.data, self.kmer_len, table_ptr,
                             self.seq_starts.ctypes.data, C.cast(hdrs, C.c_void_p), len(self.headers))
        if rc:
            raise RuntimeError(L.fx_last_error().decode())
        return path


def concat_records(records, alphabet):
    """libsufr's read_sequence_file model: records upper-cased and joined by 'N'/'X'
    (fm_index.rs:148-153).  -> (text uint8[], starts uint64[])"""
    delim = b"N" if alphabet == NUCLEOTIDE else b"X"
    recs = [r.encode() if isinstance(r, str) else bytes(r) for r in records]
    starts, pos = [], 0
    for r in recs:
        starts.append(pos)
        pos += len(r) + 1
    text = np.frombuffer(delim.join(recs).upper(), dtype=np.uint8).copy()
    return text, np.array(starts, dtype=np.uint64)


def build_parts(text, alphabet=NUCLEOTIDE, ratio=8, kmer_len=None, seq_starts=None, headers=None,
                keep_sa=False):
    """text: uint8 array / bytes WITHOUT the trailing '$' (it is implied)."""
    L = lib()
    if isinstance(text, (bytes, str)):
        text = np.frombuffer(text.encode() if isinstance(text, str) else text, dtype=np.uint8)
    text = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(text)
    bwt_len = n + 1
    if kmer_len is None:
        kmer_len = 10 if alphabet == NUCLEOTIDE else 4  # kmer_lookup_table.rs:23-24
    sa = np.empty(bwt_len, dtype=np.uint32)
    if L.fx_suffix_array(alphabet, text.ctypes.data, n, sa.ctypes.data):
        raise RuntimeError(L.fx_last_error().decode())
    blocks = np.zeros(L.fx_num_blocks(bwt_len) * L.fx_block_words(alphabet), dtype=np.uint64)
    card = 6 if alphabet == NUCLEOTIDE else 22
    prefix_sums = np.zeros(card + 1, dtype=np.uint64)
    sa_words = np.zeros(L.fx_sa_words(bwt_len, ratio) + 1, dtype=np.uint64)[:L.fx_sa_words(bwt_len, ratio)]
    if L.fx_build_parts(alphabet, text.ctypes.data, n, sa.ctypes.data, ratio, blocks.ctypes.data,
                        prefix_sums.ctypes.data, sa_words.ctypes.data):
        raise RuntimeError(L.fx_last_error().decode())
    if seq_starts is None:
        seq_starts = np.zeros(1, dtype=np.uint64)
    if headers is None:
        headers = ["synthetic"] if len(seq_starts) == 1 else [f"seq{i}" for i in range(len(seq_starts))]
    return Parts(alphabet, ratio, bwt_len, kmer_len, blocks, prefix_sums, sa_words,
                 np.ascontiguousarray(seq_starts, dtype=np.uint64), headers, text,
                 sa if keep_sa else None)
