"""ncu target for the protein search kernel (BASELINE cfg4): 10 M x 12-residue peptides vs 2 G residues, k = 5."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from awry_b200 import FmIndex
from fixtures import pyfixture_gpu as fxg
n, nq, L = 2_000_000_000, 10_000_000, 12
parts, _ = fxg.build_parts(1, n, 6, ratio=8, kmer_len=5)
ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks, parts.prefix_sums, parts.sa_words)
d = torch.empty(nq * L, dtype=torch.uint8, device="cuda"); fxg.gen_queries_device(1, n, 6, nq, L, 7, d.data_ptr())
off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * L
cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    ix.count_device(d.data_ptr(), off.data_ptr(), nq, cnt.data_ptr(), st)
torch.cuda.synchronize()
print("ok", int(cnt.min()))
