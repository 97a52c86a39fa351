#!/bin/bash
# A/B of the tau kernel's occupancy builds + parity suite + ncu capture of the fastest
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tau.py tests/test_gpu_direct.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
best=3; bestms=1e9
for occ in 3 4 5; do
  VGSIM_TAU_OCC=$occ VGSIM_BENCH_WATCHDOG=150 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_occ$occ.log 2> gpurun_out/bench_occ$occ.err
  ms=$(python -c "import json;print(json.loads(open('gpurun_out/bench_occ$occ.log').read().strip().splitlines()[-1])['roofline']['kernel_ms'])" 2>/dev/null || echo 1e9)
  echo "occ $occ kernel_ms $ms" >> gpurun_out/occ_summary.txt
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)"; then best=$occ; bestms=$ms; fi
done
echo "best $best $bestms" >> gpurun_out/occ_summary.txt
VGSIM_TAU_OCC=$best timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_kernel -s 1 -c 1 -o gpurun_out/prof_tau \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --replicates 1776 --leaps 16 > gpurun_out/bench_ncu_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/occ_summary.txt
