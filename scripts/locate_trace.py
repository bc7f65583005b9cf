import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n=3_100_000_000
parts,_=fxg.build_parts(0,n,3,ratio=8,kmer_len=13)
ix=FmIndex.from_parts(parts.alphabet,parts.ratio,parts.bwt_len,parts.kmer_len,parts.blocks,parts.prefix_sums,parts.sa_words)
nl,ll=1_000_000,50
d=torch.empty(nl*ll,dtype=torch.uint8,device="cuda"); fxg.gen_queries_device(0,n,3,nl,ll,5,d.data_ptr())
hq=torch.empty(nl*ll,dtype=torch.uint8,pin_memory=True); hq.copy_(d)
ho=torch.empty(nl+1,dtype=torch.int64,pin_memory=True); ho.copy_(torch.arange(0,nl+1,dtype=torch.int64)*ll)
qb,qo=hq.numpy(),ho.numpy().view(np.uint64)
hoff=torch.zeros(nl+1,dtype=torch.int64,pin_memory=True).numpy().view(np.uint64)
hits=torch.zeros((nl+1024,2),dtype=torch.int64,pin_memory=True).numpy().view(np.uint64)
for i in range(4):
    if i==3: print("---- traced call", file=sys.stderr)
    t0=time.perf_counter(); ix.locate_packed_into(qb,qo,hoff,hits); dt=time.perf_counter()-t0
    print(f"call {i}: {dt*1e3:.2f} ms", file=sys.stderr)
