#!/bin/bash
# Development check in ONE gpurun call: a subset (or all) of the GPU tests, then the bench line.
#   gpurun --timeout 600 -- 'bash scripts/r2_check.sh <tag> "<pytest args>" "<bench args>"'
T=${1:-r02}
PT=${2:-tests -m gpu -x -q}
BA=${3:-}
mkdir -p gpurun_out
nproc > gpurun_out/${T}_host.txt; nvidia-smi -L >> gpurun_out/${T}_host.txt; free -g | head -2 >> gpurun_out/${T}_host.txt
timeout 420 python -m pytest $PT > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/${T}_tests.log
if [ "$BA" != "skip" ]; then
  timeout 300 python bench.py $BA > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
  tail -5 gpurun_out/${T}_bench.err
  cut -c1-1500 gpurun_out/${T}_bench.json
fi
