"""Does host-side packing (16 threads streaming through host DRAM) slow a GPU kernel that never touches host
memory?  search of 512 k x 50-bp device-resident queries, CUDA events, under different host loads."""
import os, sys, time, threading
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n = 3_100_000_000
parts, _ = fxg.build_parts(0, n, 3, ratio=8, kmer_len=13)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
nl, ll = 1 << 19, 50
d = torch.empty(nl * ll, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(0, n, 3, nl, ll, 5, d.data_ptr())
off = torch.arange(0, nl + 1, dtype=torch.int64, device="cuda") * ll
cnt = torch.zeros(nl, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
src = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, 64 << 20)].copy()
stop = False
def packer():
    while not stop:
        f.host_pack_dna(src)
def spinner():
    x = 0
    a = np.arange(4096, dtype=np.int64)
    while not stop:
        x += int(a.sum())
def measure(tag):
    f.profile_enable(True)
    ts = []
    for _ in range(30):
        f.profile_reset()
        ix.count_device(d.data_ptr(), off.data_ptr(), nl, cnt.data_ptr(), st)
        torch.cuda.synchronize()
        p = f.profile_get()
        ts.append(p["search_ms"])
    f.profile_enable(False)
    ts.sort()
    print(f"{tag}: search kernel min {ts[0]:.3f} median {ts[len(ts)//2]:.3f} max {ts[-1]:.3f} ms", flush=True)
measure("host idle")
for threads, name in ((16, "16 packer threads"), (8, "8 packer threads"), (2, "2 packer threads")):
    f.set_host_threads(threads)
    stop = False
    th = threading.Thread(target=packer); th.start()
    time.sleep(0.2)
    measure(name)
    stop = True; th.join()
f.set_host_threads(0)
stop = False
ths = [threading.Thread(target=spinner) for _ in range(4)]
[t.start() for t in ths]
time.sleep(0.2)
measure("4 python spinner threads (GIL-bound)")
stop = True
[t.join() for t in ths]
measure("host idle again")
