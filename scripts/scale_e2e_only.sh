#!/bin/bash
# e2e leg only, at the GPU counts given as arguments, with the mixed pack/raw schedule on and off
for mixed in 1 0; do
for n in "$@"; do
  AWRY_B200_PACK_MIXED=$mixed python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 --no-locate --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed=$mixed n=$n e2e %.1f M reads/s  ms %.2f  h2d MB %d' % (d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']//1000000))"
done; done
