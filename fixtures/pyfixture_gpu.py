"""GPU-side test/bench fixtures: synthetic text and queries generated on the device
(fixtures/libawry_fixture_gpu.so), and index parts at BASELINE scale.

The index itself is built by the PRODUCT's GPU builder (awry_b200.fm_index.build_parts ->
awry_build_parts), which tests/test_gpu_fixture.py checks bit for bit against the independent CPU
builder of fixtures/fixture_cpu.cpp.  Test/bench infrastructure only."""
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
        L.fxg_gen_text_device.argtypes = [C.c_int, u64, u64, vp, vp]
        L.fxg_gen_queries_device.argtypes = [C.c_int, u64, u64, u64, u64, u64, C.c_uint32, vp, vp]
        _LIB = L
    return _LIB


def build_parts(alphabet, n, text_seed, ratio=8, kmer_len=None, device=0, host_text=None):
    """-> (fixtures.pyfixture.Parts, phase seconds dict).  The text is the synthetic one of
    (alphabet, text_seed), generated on the device, unless host_text (uint8 ASCII, n bytes) is given."""
    import torch
    from awry_b200 import fm_index as f
    if kmer_len is None:
        kmer_len = 10 if alphabet == 0 else 4
    if host_text is not None:
        host_text = np.ascontiguousarray(host_text, dtype=np.uint8)
        assert len(host_text) == n
        blocks, prefix, sa_words, phases = f.build_parts(alphabet, host_text, sa_ratio=ratio, device=device)
    else:
        with torch.cuda.device(device):
            d_text = torch.empty(n, dtype=torch.uint8, device="cuda")
            if lib().fxg_gen_text_device(alphabet, n, text_seed, d_text.data_ptr(), 0):
                raise RuntimeError(lib().fxg_last_error().decode())
            torch.cuda.synchronize()
            blocks, prefix, sa_words, phases = f.build_parts(alphabet, d_text.data_ptr(), n=n, sa_ratio=ratio,
                                                             device=device)
            del d_text
    parts = cpu.Parts(alphabet, ratio, n + 1, kmer_len, blocks, prefix, sa_words,
                      np.zeros(1, dtype=np.uint64), ["synthetic"], host_text)
    return parts, phases


def gen_queries_device(alphabet, n, text_seed, nq, qlen, qseed, d_out_ptr, mut_ppm=0, stream=0):
    rc = lib().fxg_gen_queries_device(alphabet, n, text_seed, nq, qlen, qseed, mut_ppm, d_out_ptr, stream)
    if rc:
        raise RuntimeError(lib().fxg_last_error().decode())
