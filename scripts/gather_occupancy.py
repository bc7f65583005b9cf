"""Random 128-B gather rate (4 lanes x LDG.256 per read, 4 GiB footprint) vs resident blocks per SM,
independent reads in flight per lane group, and static vs dynamic distribution of the work over the SMs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402

print("static split: one wave of blocks, equal work per SM.  G reads/s")
for bps in (2, 4, 8):
    row = []
    for un in (1, 2, 4, 8):
        r, g = f.bench_random_gather(0, 4 << 30, 128, 1000 + 10 * bps + un, 200_000_000, 2)
        row.append(f"unroll {un}: {r/1e9:5.1f}")
    print(f"  {bps} blocks/SM:  " + "   ".join(row), flush=True)
print("dynamic: W x as many 256-thread blocks as are resident (8/SM), scheduled by the hardware")
for waves in (1, 4, 16, 64):
    row = []
    for un in (1, 2, 4, 8):
        r, g = f.bench_random_gather(0, 4 << 30, 128, 3000 + 10 * waves + un, 200_000_000, 2)
        row.append(f"unroll {un}: {r/1e9:5.1f}")
    print(f"  waves {waves:2d}:  " + "   ".join(row), flush=True)
