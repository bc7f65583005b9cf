#!/bin/bash
# state-machine pair kernel: parity (whole GPU suite under AWRY_B200_SLOTS=2 and =1) and A/B timing
T=${1:-r02k}
mkdir -p gpurun_out
for S in 2 1; do
  AWRY_B200_SLOTS=$S timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_round2.py tests/test_gpu_reads_file.py -m gpu -x -q > gpurun_out/${T}_tests_slots$S.log 2>&1; echo "slots=$S tests rc=$?"; tail -2 gpurun_out/${T}_tests_slots$S.log
done
V="80:6,80:8,81:6,81:8,81:4,82:4,82:3,82:5"
timeout 200 python scripts/ab_search.py --variants $V --reps 4 --burst 10 > gpurun_out/${T}_ab_exact.log 2>&1; cat gpurun_out/${T}_ab_exact.log | tail -9
timeout 200 python scripts/ab_search.py --variants $V --reps 4 --burst 10 --mut-ppm 100000 > gpurun_out/${T}_ab_mut10.log 2>&1; cat gpurun_out/${T}_ab_mut10.log | tail -9
timeout 100 python scripts/ab_search.py --variants $V --reps 4 --burst 10 --nq 1000000 --qlen 50 > gpurun_out/${T}_ab_50bp.log 2>&1; cat gpurun_out/${T}_ab_50bp.log | tail -9
