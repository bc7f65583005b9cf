"""Host-side 2-bit query packer (awry_b200/csrc/hostpack.cpp): every SIMD level against a numpy
restatement.  Pure CPU; the device half (pack2_kernel) is covered by the GPU parity tests, which run the
batch entry points with host packing both on and off."""
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
rng = np.random.default_rng(7)
for n in (0, 1, 3, 4, 63, 64, 65, 127, 1000, 4097, 300_001, 3_000_000):
    src = np.frombuffer(b"ACGTacgt", dtype=np.uint8)[rng.integers(0, 8, n)].copy()
    k = max(1, n // 50)
    if n:
        pos = rng.integers(0, n, k)
        odd = np.frombuffer(b"NnRYKM$#U-*X\n \x00\x01\x10\x21\x41\x80\xc1\xe7\xd4\xff", dtype=np.uint8)
        src[pos] = odd[rng.integers(0, len(odd), k)]       # incl. NUL / space / bytes with bit 7 set
    got, exc = f.host_pack_dna(src)
    up = src & 0xDF
    bad = ~((up == 65) | (up == 67) | (up == 71) | (up == 84))
    want_exc = (np.nonzero(bad)[0].astype(np.uint64) << np.uint64(8)) | src[bad].astype(np.uint64)
    assert np.array_equal(np.sort(exc), want_exc), (n, len(exc), len(want_exc))
    crumbs = ((src >> 1) & 3).astype(np.uint8)
    pad = np.zeros((-n) %% 4, dtype=np.uint8)
    c4 = np.concatenate([crumbs, pad]).reshape(-1, 4)
    want = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    assert np.array_equal(got, want), n
print("ok")
""" % ROOT


@pytest.mark.parametrize("simd", ["0", "1", "2"])
def test_host_packer_matches_numpy(simd):
    env = dict(os.environ, AWRY_B200_HOST_SIMD=simd, AWRY_B200_HOST_THREADS="5")
    out = subprocess.run([sys.executable, "-c", CHECK], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]


def test_exceptions_come_back_ascending():
    """awry_*_batch_packed2 wants the exception list sorted by position: the packer's own output is"""
    from awry_b200 import fm_index as f
    rng = np.random.default_rng(3)
    src = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 2_000_003)].copy()
    src[rng.integers(0, len(src), 5000)] = ord("N")
    _, exc = f.host_pack_dna(src)
    pos = exc >> np.uint64(8)
    assert len(exc) > 4000 and np.all(pos[1:] > pos[:-1])


def test_pool_serves_concurrent_callers_and_resizes():
    """one process-wide pool, several parallel regions open at once (one per replica thread of a multi-GPU
    call): every caller gets its own, complete result, whatever the thread cap"""
    import threading
    from awry_b200 import fm_index as f
    rng = np.random.default_rng(11)
    srcs = [np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.integers(0, 9, 1_500_000 + 77_777 * i)].copy() for i in range(6)]
    want = [f.host_pack_dna(s) for s in srcs]
    before = f.host_threads()
    try:
        for cap in (1, 3, 8):
            f.set_host_threads(cap)
            assert f.host_threads() == cap
            got, errs = [None] * len(srcs), []

            def work(i):
                try:
                    for _ in range(3):
                        got[i] = f.host_pack_dna(srcs[i])
                except Exception as e:  # noqa: BLE001
                    errs.append(e)

            th = [threading.Thread(target=work, args=(i,)) for i in range(len(srcs))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not errs, errs
            for (a, ea), (b, eb) in zip(want, got):
                assert np.array_equal(a, b) and np.array_equal(ea, eb)
    finally:
        f.set_host_threads(0)
    assert f.host_threads() == before
