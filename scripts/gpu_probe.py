"""First-contact probe for a gpurun box: host shape, random-gather roofline sweep, and a
small-index timing of the search-kernel variants (L2-resident: plumbing check, not a bench)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from awry_b200 import fm_index as f  # noqa: E402


def main():
    out = {"nproc": os.cpu_count()}
    try:
        with open("/proc/meminfo") as fh:
            out["mem_total_gb"] = int(fh.readline().split()[1]) / 1e6
    except Exception:
        pass
    import torch
    out["gpu"] = torch.cuda.get_device_name(0)
    out["n_gpus"] = torch.cuda.device_count()
    sweep = []
    fp = 4 << 30
    for granule, lanes in [(32, 1), (32, 2), (64, 1), (64, 2), (64, 4), (128, 1), (128, 2), (128, 4), (128, 8)]:
        r, g = f.bench_random_gather(0, fp, granule, lanes, 400_000_000, 3)
        sweep.append({"granule": granule, "lanes": lanes, "reads_per_s": r, "gb_per_s": g})
        print(f"gather granule={granule:4d} lanes={lanes} : {r/1e9:7.2f} G reads/s  {g:8.1f} GB/s", flush=True)
    out["gather_4GiB"] = sweep
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.json", "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
