"""two configurations of the 128-B random-gather probe for an `ncu --set full` comparison"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f
for code in [int(x) for x in (sys.argv[1:] or ["1084", "1044"])]:
    r, g = f.bench_random_gather(0, 4 << 30, 128, code, 200_000_000, 0)
    print(code, r / 1e9, flush=True)
