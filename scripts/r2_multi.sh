#!/bin/bash
# multi-GPU check in ONE gpurun --gpus N call: the in-process multi-replica tests, then bench.py under torchrun
N=${1:-2}
T=${2:-r02m}
mkdir -p gpurun_out
nproc > gpurun_out/${T}_host.txt; nvidia-smi -L >> gpurun_out/${T}_host.txt; free -g | head -2 >> gpurun_out/${T}_host.txt
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/${T}_host.txt
nvidia-smi topo -m >> gpurun_out/${T}_host.txt 2>&1
timeout 300 python -m pytest "tests/test_gpu_round2.py::test_one_batch_over_several_replicas" "tests/test_gpu_parity.py::test_multi_replica_in_one_process" "tests/test_gpu_reads_file.py::test_reads_file_over_several_replicas" -m gpu -x -q -rs > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${T}_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_bench_n$N.json") if l.startswith("{")][-1])
    for k in ("value", "ms_per_step", "e2e", "e2e_prepacked", "e2e_per_process", "e2e_prepacked_per_process", "inprocess_replicas"):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print("no bench line", e)
PY
