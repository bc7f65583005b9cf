"""Memory-safety evidence without compute-sanitizer (closed on this GPU pool: profiles/r02_s3_compute_sanitizer_refused.log).
`make -C awry_b200/csrc checked` builds libawry_b200_checked.so: the same library with every load whose address is
computed from index or query data bounds-checked against the array sizes (AWRY_CHK, layout.cuh); a violation is a
device assert.  The GPU parity suites run against that build in a subprocess (AWRY_B200_LIB), and a second
subprocess shows the checks are live by searching an index whose prefix sums contradict its BWT."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
CHECKED = os.path.join(ROOT, "awry_b200", "libawry_b200_checked.so")


@pytest.fixture(scope="module")
def checked_lib():
    if not os.path.exists(CHECKED):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "awry_b200", "csrc"), "checked"])
    return CHECKED


def test_parity_suites_pass_on_the_bounds_checked_build(checked_lib):
    env = dict(os.environ, AWRY_B200_LIB=checked_lib)
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "tests/test_gpu_fuzz.py",
                          "tests/test_gpu_round2.py", "tests/test_gpu_wide.py", "tests/test_gpu_text_finish.py", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
                          "-k", "not 4p6 and not several_replicas and not multi_replica and not cxx_host_mirror"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = out.stdout[-1500:] + out.stderr[-1500:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "Assertion" not in out.stderr, tail


LIVE = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture as fx
assert f.library_path().endswith("_checked.so")
text = fx.gen_text(0, 50_000, 3)
p = fx.build_parts(text, 0, ratio=8, kmer_len=0)
n = p.bwt_len
lie = np.array([0, 1, n, n, n, n, n], dtype=np.uint64)      # "every symbol is A": C rows then point past the BWT
print("loading", flush=True)
ix = FmIndex.from_parts(p.alphabet, p.ratio, n, 0, p.blocks, lie, p.sa_words)    # (the pair-index build already follows LF)
qb, qo = f.pack_queries([b"ACGTACGTACGTACGTACGT", b"CCCCCCCC", b"GATTACA"])
print("searching", flush=True)
print("COUNTS", ix.count_packed(qb, qo))
"""


def test_the_checks_are_live(checked_lib):
    """an index whose prefix sums lie about its BWT sends row pointers past the rank blocks: the checked build
    stops at the first such load with a device assert instead of reading whatever lies behind the array"""
    env = dict(os.environ, AWRY_B200_LIB=checked_lib, AWRY_B200_KMER_DEV="0")
    out = subprocess.run([sys.executable, "-c", LIVE % ROOT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    both = out.stdout[-800:] + out.stderr[-2500:]
    assert "loading" in out.stdout and "COUNTS" not in out.stdout, both
    assert out.returncode != 0 and "Assertion" in out.stderr and "device-side assert" in out.stderr, both
