"""Regenerates tests/golden/appendix_a.awry: the worked example of SURVEY.md Appendix A
(text GATTACAGATTACANACGT, ratio 4, k 2), written by the fixture builder in the reference's
`.awry` v1 format.  Its sha256 is pinned in tests/test_oracle.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fixtures import pyfixture as fx  # noqa: E402

if __name__ == "__main__":
    parts = fx.build_parts(b"GATTACAGATTACANACGT", fx.NUCLEOTIDE, ratio=4, kmer_len=2)
    parts.write(os.path.join(os.path.dirname(os.path.abspath(__file__)), "appendix_a.awry"))
