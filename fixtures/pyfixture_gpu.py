"""ctypes binding of the GPU fixture builder (fixtures/libawry_fixture_gpu.so): synthetic text ->
reference-layout index parts at BASELINE scale, and device-side synthetic queries.
Test/bench infrastructure only -- the product package never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import pyfixture as cpu

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "gpu"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libawry_fixture_gpu.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, u64 = C.c_void_p, C.c_uint64
        L.fxg_last_error.restype = C.c_char_p
        L.fxg_build.argtypes = [C.c_int, u64, u64, vp, u64, C.c_int, vp, vp, vp, vp]
        L.fxg_gen_queries_device.argtypes = [C.c_int, u64, u64, u64, u64, u64, C.c_uint32, vp, vp]
        _LIB = L
    return _LIB


def build_parts(alphabet, n, text_seed, ratio=8, kmer_len=None, device=0, host_text=None):
    """-> (fixtures.pyfixture.Parts with text=None, phase seconds dict).  host_text (uint8 ASCII,
    n bytes) overrides the synthetic text of (alphabet, text_seed)."""
    L = lib()
    c = cpu.lib()
    bwt_len = n + 1
    if kmer_len is None:
        kmer_len = 10 if alphabet == 0 else 4
    blocks = np.empty(c.fx_num_blocks(bwt_len) * c.fx_block_words(alphabet), dtype=np.uint64)
    card = 6 if alphabet == 0 else 22
    prefix_sums = np.zeros(card + 1, dtype=np.uint64)
    sa_words = np.empty(c.fx_sa_words(bwt_len, ratio), dtype=np.uint64)
    phases = np.zeros(8, dtype=np.float64)
    tptr = None
    if host_text is not None:
        host_text = np.ascontiguousarray(host_text, dtype=np.uint8)
        assert len(host_text) == n
        tptr = host_text.ctypes.data
    rc = L.fxg_build(alphabet, n, text_seed, tptr, ratio, device, blocks.ctypes.data,
                     prefix_sums.ctypes.data, sa_words.ctypes.data, phases.ctypes.data)
    if rc:
        raise RuntimeError(L.fxg_last_error().decode())
    names = ["gen", "keys", "sort", "ties", "bwt", "milestones", "sa_pack_copy", "total"]
    parts = cpu.Parts(alphabet, ratio, bwt_len, kmer_len, blocks, prefix_sums, sa_words,
                      np.zeros(1, dtype=np.uint64), ["synthetic"], host_text)
    return parts, dict(zip(names, [float(x) for x in phases]))


def gen_queries_device(alphabet, n, text_seed, nq, qlen, qseed, d_out_ptr, mut_ppm=0, stream=0):
    rc = lib().fxg_gen_queries_device(alphabet, n, text_seed, nq, qlen, qseed, mut_ppm, d_out_ptr, stream)
    if rc:
        raise RuntimeError(lib().fxg_last_error().decode())
