"""ctypes binding of the CPU oracle (oracle/libawry_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package awry_b200 never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("lf_steps", "block_touches", "seeded_steps", "seeded_touches", "walk_steps", "hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Hit(C.Structure):
    _fields_ = [("seq_idx", C.c_uint64), ("local_pos", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libawry_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, u64, u8p, u64p = C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p
        L.awo_last_error.restype = C.c_char_p
        L.awo_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.awo_from_parts.argtypes = [C.c_int, u64, u64, C.c_uint, vp, vp, vp, vp, u64, C.POINTER(vp)]
        L.awo_free.argtypes = [vp]
        L.awo_free_ptr.argtypes = [vp]
        L.awo_ascii_to_index.argtypes = [C.c_int, C.c_uint8]
        L.awo_ascii_to_index.restype = C.c_uint8
        L.awo_index_to_code.argtypes = [C.c_int, C.c_uint8]
        L.awo_index_to_code.restype = C.c_uint8
        L.awo_code_to_index.argtypes = [C.c_int, C.c_uint8]
        L.awo_code_to_index.restype = C.c_uint8
        L.awo_index_to_ascii.argtypes = [C.c_int, C.c_uint8]
        L.awo_index_to_ascii.restype = C.c_char
        L.awo_bits_per_element.argtypes = [u64]
        L.awo_bits_per_element.restype = C.c_uint
        L.awo_compressed_word_len.argtypes = [u64, u64]
        L.awo_compressed_word_len.restype = u64
        L.awo_sa_set_value.argtypes = [vp, C.c_uint, u64, u64]
        L.awo_sa_reconstruct.argtypes = [vp, u64]
        L.awo_sa_reconstruct.restype = u64
        L.awo_masked_popcount.argtypes = [vp, u64]
        L.awo_masked_popcount.restype = C.c_uint32
        L.awo_block_occurrence.argtypes = [vp, vp, u64, C.c_uint8]
        L.awo_block_occurrence.restype = u64
        L.awo_global_occurrence.argtypes = [vp, u64, C.c_uint8]
        L.awo_global_occurrence.restype = u64
        L.awo_symbol_at.argtypes = [vp, u64]
        L.awo_symbol_at.restype = C.c_uint8
        L.awo_initial_range.argtypes = [vp, C.c_uint8, C.POINTER(u64), C.POINTER(u64)]
        L.awo_update_range.argtypes = [vp, u64, u64, C.c_uint8, C.POINTER(u64), C.POINTER(u64)]
        L.awo_backstep.argtypes = [vp, u64]
        L.awo_backstep.restype = u64
        L.awo_search_range.argtypes = [vp, u8p, u64, C.POINTER(u64), C.POINTER(u64), vp]
        L.awo_count_string.argtypes = [vp, u8p, u64, C.POINTER(u64)]
        L.awo_locate_string.argtypes = [vp, u8p, u64, C.POINTER(vp), C.POINTER(u64), vp]
        L.awo_count_batch.argtypes = [vp, u8p, u64p, u64, u64p, C.c_int, C.POINTER(Stats)]
        L.awo_locate_batch.argtypes = [vp, u8p, u64p, u64, u64p, C.POINTER(vp), C.POINTER(u64),
                                       C.c_int, C.c_int, C.POINTER(Stats)]
        L.awo_hw_threads.restype = C.c_int
        _LIB = L
    return _LIB


class _IndexStruct(C.Structure):
    _fields_ = [("version", C.c_uint64), ("sa_ratio", C.c_uint64), ("bwt_len", C.c_uint64),
                ("alphabet", C.c_int), ("card", C.c_int), ("n_planes", C.c_int),
                ("n_milestones", C.c_int), ("block_words", C.c_size_t), ("n_blocks", C.c_uint64),
                ("blocks", C.c_void_p), ("prefix_sums", C.c_uint64 * 23), ("sa_words", C.c_void_p),
                ("n_sa_words", C.c_uint64), ("bits", C.c_uint), ("kmer_len", C.c_uint),
                ("n_seqs", C.c_uint64), ("seq_starts", C.c_void_p), ("headers", C.c_void_p),
                ("owns_arrays", C.c_int)]


def pack_queries(queries):
    """list of bytes/str -> (uint8 array, uint64 offsets[nq+1])"""
    qs = [q.encode() if isinstance(q, str) else bytes(q) for q in queries]
    off = np.zeros(len(qs) + 1, dtype=np.uint64)
    if qs:
        off[1:] = np.cumsum([len(q) for q in qs], dtype=np.uint64)
    data = np.frombuffer(b"".join(qs), dtype=np.uint8).copy() if qs else np.zeros(0, np.uint8)
    return data, off


class OracleError(RuntimeError):
    pass


class OracleIndex:
    """Restatement of awry::fm_index::FmIndex (query side) on the CPU."""

    def __init__(self, handle, keep=None):
        self._h = C.c_void_p(handle)
        self._keep = keep
        self._s = C.cast(self._h, C.POINTER(_IndexStruct)).contents

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        if lib().awo_load(os.fsencode(path), C.byref(h)):
            raise OracleError(lib().awo_last_error().decode())
        return cls(h.value)

    @classmethod
    def from_parts(cls, alphabet, sa_ratio, bwt_len, kmer_len, blocks, prefix_sums, sa_words,
                   seq_starts=None):
        if seq_starts is None:
            seq_starts = np.zeros(1, np.uint64)
        arrays = [np.ascontiguousarray(a, dtype=np.uint64) for a in
                  (blocks, prefix_sums, sa_words, seq_starts)]
        h = C.c_void_p()
        rc = lib().awo_from_parts(alphabet, sa_ratio, bwt_len, kmer_len, arrays[0].ctypes.data,
                                  arrays[1].ctypes.data, arrays[2].ctypes.data,
                                  arrays[3].ctypes.data, len(arrays[3]), C.byref(h))
        if rc:
            raise OracleError(lib().awo_last_error().decode())
        return cls(h.value, keep=arrays)

    def close(self):
        if self._h:
            lib().awo_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (fm_index.rs:302-368)
    alphabet = property(lambda s: s._s.alphabet)
    bwt_len = property(lambda s: s._s.bwt_len)
    sa_ratio = property(lambda s: s._s.sa_ratio)
    kmer_len = property(lambda s: s._s.kmer_len)
    version = property(lambda s: s._s.version)
    bits = property(lambda s: s._s.bits)
    n_blocks = property(lambda s: s._s.n_blocks)
    n_seqs = property(lambda s: s._s.n_seqs)

    @property
    def prefix_sums(self):
        return [int(self._s.prefix_sums[i]) for i in range(self._s.card + 1)]

    def sym(self, ch):
        if isinstance(ch, str):
            ch = ord(ch)
        return lib().awo_ascii_to_index(self.alphabet, ch)

    def global_occurrence(self, pos, sym):
        return lib().awo_global_occurrence(self._h, pos, sym)

    def symbol_at(self, pos):
        return lib().awo_symbol_at(self._h, pos)

    def initial_range(self, sym):
        sp, ep = C.c_uint64(), C.c_uint64()
        lib().awo_initial_range(self._h, sym, C.byref(sp), C.byref(ep))
        return sp.value, ep.value

    def update_range(self, sp, ep, sym):
        a, b = C.c_uint64(), C.c_uint64()
        lib().awo_update_range(self._h, sp, ep, sym, C.byref(a), C.byref(b))
        return a.value, b.value

    def backstep(self, pos):
        return lib().awo_backstep(self._h, pos)

    def sa_reconstruct(self, row):
        return lib().awo_sa_reconstruct(self._h, row)

    def search_range(self, q):
        q = q.encode() if isinstance(q, str) else bytes(q)
        sp, ep = C.c_uint64(), C.c_uint64()
        buf = C.create_string_buffer(q, len(q))
        if lib().awo_search_range(self._h, C.cast(buf, C.c_void_p), len(q), C.byref(sp), C.byref(ep), None):
            raise OracleError(lib().awo_last_error().decode())
        return sp.value, ep.value

    def count_string(self, q):
        sp, ep = self.search_range(q)
        return 0 if sp > ep else ep - sp + 1

    def locate_string(self, q):
        q = q.encode() if isinstance(q, str) else bytes(q)
        hits, n = C.c_void_p(), C.c_uint64()
        buf = C.create_string_buffer(q, len(q))
        if lib().awo_locate_string(self._h, C.cast(buf, C.c_void_p), len(q), C.byref(hits), C.byref(n), None):
            raise OracleError(lib().awo_last_error().decode())
        out = []
        if n.value:
            arr = np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_uint64)), shape=(n.value, 2)).copy()
            out = [(int(a), int(b)) for a, b in arr]
        lib().awo_free_ptr(hits)
        return out

    def count_batch(self, qbytes, qoff, n_threads=0):
        """parallel_count on packed queries -> (counts uint64[nq], stats dict)"""
        qbytes = np.ascontiguousarray(qbytes, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        nq = len(qoff) - 1
        counts = np.zeros(nq, dtype=np.uint64)
        st = Stats()
        if lib().awo_count_batch(self._h, qbytes.ctypes.data, qoff.ctypes.data, nq,
                                 counts.ctypes.data, n_threads, C.byref(st)):
            raise OracleError(lib().awo_last_error().decode())
        return counts, st.as_dict()

    def locate_batch(self, qbytes, qoff, sorted_hits=False, n_threads=0):
        """parallel_locate -> (hit_off uint64[nq+1], hits uint64[n,2], stats dict)"""
        qbytes = np.ascontiguousarray(qbytes, dtype=np.uint8)
        qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
        nq = len(qoff) - 1
        hit_off = np.zeros(nq + 1, dtype=np.uint64)
        hits, n = C.c_void_p(), C.c_uint64()
        st = Stats()
        if lib().awo_locate_batch(self._h, qbytes.ctypes.data, qoff.ctypes.data, nq,
                                  hit_off.ctypes.data, C.byref(hits), C.byref(n),
                                  1 if sorted_hits else 0, n_threads, C.byref(st)):
            raise OracleError(lib().awo_last_error().decode())
        arr = np.empty((n.value, 2), dtype=np.uint64)
        if n.value:
            C.memmove(arr.ctypes.data, hits, n.value * 16)
        lib().awo_free_ptr(hits)
        return hit_off, arr, st.as_dict()
