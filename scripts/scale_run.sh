#!/bin/bash
# bench.py at 1/2/4/8 GPUs of one box, the way the driver launches it (torchrun for N > 1)
set -u
nproc; nvidia-smi -L | wc -l
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r01b_scale_n$n.json 2> gpurun_out/r01b_scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r01b_scale_n$n.json 2> gpurun_out/r01b_scale_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r01b_scale_n$n.json").read().strip().splitlines()[-1])
    print("n=$n value %.1f M reads/s  ms/step %.2f  e2e %.1f M reads/s (%s)  locate %.2f G hits/s" % (
        d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["host_pack"][:20], d["locate"]["hits_per_s"] / 1e9))
except Exception as e:
    print("n=$n failed", e)
PY
done
