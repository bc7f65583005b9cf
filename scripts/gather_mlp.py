"""Random 128-B gather rate vs requests in flight and issue path (LDG.256 x 4 lanes vs TMA bulk copy)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402

names = {4: "LDG.256 x4 lanes, 4 in flight/group", 102: "LDG.256 x4 lanes, 2 in flight/group",
         104: "LDG.256 x4 lanes, 1 in flight/group", 201: "cp.async.bulk 128 B per thread, 1 in flight/thread",
         202: "cp.async.bulk 128 B per thread, 2 in flight/thread"}
for lanes in (4, 102, 104, 201, 202):
    r, g = f.bench_random_gather(0, 4 << 30, 128, lanes, 400_000_000, 2)
    print(f"{names[lanes]:52s}: {r/1e9:6.2f} G reads/s  {g:7.1f} GB/s", flush=True)
