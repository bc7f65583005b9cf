import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from awry_b200 import FmIndex, fm_index as f
from fixtures import pyfixture_gpu as fxg
n=3_100_000_000
parts,_=fxg.build_parts(0,n,3,ratio=8,kmer_len=13)
os.environ["AWRY_B200_FULL_SA"]="0"
ix=FmIndex.from_parts(parts.alphabet,parts.ratio,parts.bwt_len,parts.kmer_len,parts.blocks,parts.prefix_sums,parts.sa_words)
nq,L=10_000_000,150
st=torch.cuda.current_stream().cuda_stream
d_off=torch.arange(0,nq+1,dtype=torch.int64,device="cuda")*L
d_cnt=torch.zeros(nq,dtype=torch.int64,device="cuda")
f.profile_enable(True)
ppms = [int(x) for x in os.environ.get("PPMS", "0,10000,100000,1000000").split(",")]
for ppm in ppms:
    d=torch.empty(nq*L,dtype=torch.uint8,device="cuda"); fxg.gen_queries_device(0,n,3,nq,L,4,d.data_ptr(),mut_ppm=ppm)
    for it in range(3):
        f.profile_reset(); ix.count_device(d.data_ptr(),d_off.data_ptr(),nq,d_cnt.data_ptr(),st); torch.cuda.synchronize(); p=f.profile_get()
    print(f"ticket {os.environ.get('AWRY_B200_TICKET','auto')} mut_ppm {ppm}: search {p['search_ms']:.2f} ms zero-count reads {int((d_cnt==0).sum())}", flush=True)
