#!/bin/bash
# compute-sanitizer memcheck on the small configurations: the new entry points of round 2
mkdir -p gpurun_out
which compute-sanitizer > gpurun_out/r02_sanitizer.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 86 --target-processes all \
  python -m pytest tests/test_gpu_round2.py tests/test_gpu_wide.py -m gpu -x -q \
  -k "not 4p6 and not several_replicas" >> gpurun_out/r02_sanitizer.log 2>&1
echo "sanitizer rc=$?" | tee -a gpurun_out/r02_sanitizer.log
tail -15 gpurun_out/r02_sanitizer.log
