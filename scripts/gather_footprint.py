"""Random-gather rate (128-B reads, 4 lanes x LDG.256) as a function of the footprint: separates the
DRAM/L2 request-rate ceiling from address-translation effects."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402

for gib in (0.25, 1, 2, 4, 8, 16, 32, 64):
    fp = int(gib * (1 << 30))
    for granule, lanes in ((128, 4), (64, 2)):
        r, g = f.bench_random_gather(0, fp, granule, lanes, 400_000_000, 2)
        print(f"footprint {gib:6.2f} GiB granule {granule:3d}: {r/1e9:6.2f} G reads/s {g:7.1f} GB/s", flush=True)
