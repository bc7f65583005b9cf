#!/bin/bash
# Round-end evidence in ONE gpurun call (one GPU): GPU tests, the default bench line (+ the reference arm's line),
# `ncu --set full` of the dominant kernel (both flavours: finish in the text / backward search only) and of the protein kernel, and the ncu launch list of a short bench run.
# Each ncu pass runs only after the same program has exited 0 without ncu.  Outputs land in gpurun_out/<tag>_*;
# scripts/ncu_summary.py turns them into profiles/.
#   gpurun --timeout 900 -- 'bash scripts/final_evidence.sh r02'
T=${1:-evidence}
mkdir -p gpurun_out
nproc > gpurun_out/${T}_host.txt; nvidia-smi -L >> gpurun_out/${T}_host.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python -m pytest tests -m gpu -x -q -rs --durations=8 > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"
tail -14 gpurun_out/${T}_tests.log
timeout 200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; rc=$?; echo "bench rc=$rc"
cut -c1-300 gpurun_out/${T}_bench.json
[ $rc -eq 0 ] || exit $rc
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference arm rc=$?"
cut -c1-300 gpurun_out/${T}_bench_reference.json
timeout 150 ncu --set full --clock-control none --import-source on -k regex:search_dna_wave_kernel -s 3 -c 1 -f \
  -o gpurun_out/${T}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-locate --no-e2e --no-secondary \
  > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:search_dna_pair_kernel -s 3 -c 1 -f \
  -o gpurun_out/${T}_prof_lf python bench.py --steps 1 --warmup 3 --count-variant 1 --no-cpu-baseline --no-locate --no-e2e --no-secondary \
  > gpurun_out/${T}_ncu_full_lf.log 2>&1; echo "ncu full (backward search only) rc=$?"
timeout 60 python scripts/cfg4_ncu_target.py > gpurun_out/${T}_cfg4_target.log 2>&1 && \
timeout 150 ncu --set full --clock-control none --import-source on -k regex:search_amino_wave_kernel -s 2 -c 1 -f \
  -o gpurun_out/${T}_prof_amino python scripts/cfg4_ncu_target.py > gpurun_out/${T}_ncu_amino.log 2>&1; echo "ncu amino rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary \
  > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python scripts/text_finish_probe.py > gpurun_out/${T}_text_finish_probe.log 2>&1; echo "probe rc=$?"
ls -la gpurun_out/${T}_*
