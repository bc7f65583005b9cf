import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# every index the GPU tests create also gets the memory-lean locate structures (walk blocks + position-sampled
# suffix array), which the library otherwise only builds when the unsampled array does not fit: the
# `locate_variant` fixtures then run all three pass-2 schemes on the same handle
os.environ.setdefault("AWRY_B200_LEAN_SA", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(path, *targets):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, path), *targets])


@pytest.fixture(scope="session", autouse=True)
def built_checkers():
    """Oracle and fixture builder are test infrastructure: build them on demand."""
    _make("oracle")
    _make("fixtures")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    return pyoracle


@pytest.fixture(scope="session")
def fx():
    from fixtures import pyfixture
    return pyfixture


def brute_positions(text: bytes, q: bytes):
    out, i = [], text.find(q)
    while i >= 0:
        out.append(i)
        i = text.find(q, i + 1)
    return out


def oracle_from_parts(po, parts):
    return po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len,
                                     parts.blocks, parts.prefix_sums, parts.sa_words, parts.seq_starts)


def device_from_parts(parts, devices=None):
    from awry_b200 import FmIndex
    return FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len,
                              parts.blocks, parts.prefix_sums, parts.sa_words, parts.seq_starts,
                              parts.headers, devices=devices)


def mixed_queries(fx, text, nq, qlen, seed, alphabet=0):
    """half exact substrings, half uniform random strings (BASELINE cfg 1)"""
    half = nq // 2
    qb, qo, _ = fx.gen_substring_queries(text, half, qlen, seed)
    rnd = fx.gen_text(alphabet, (nq - half) * qlen, seed + 1000)
    qbytes = np.concatenate([qb, rnd])
    qoff = np.arange(nq + 1, dtype=np.uint64) * np.uint64(qlen)
    return qbytes, qoff
