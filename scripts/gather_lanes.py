"""Random 128-B gather: how many lanes should share one line?  4 lanes x LDG.256, 2 lanes x 2 LDG.256,
1 lane x 4 LDG.256, all with the work balanced dynamically over the SMs (16 waves of blocks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402

for name, code in (("4 lanes x 1 LDG.256", 3164), ("2 lanes x 2 LDG.256", 4016), ("1 lane  x 4 LDG.256", 5016)):
    r, g = f.bench_random_gather(0, 4 << 30, 128, code, 200_000_000, 2)
    print(f"{name}: {r/1e9:5.1f} G reads/s", flush=True)
