#!/bin/bash
# A/B of CTA size x resident CTAs per SM for the tau kernel (instruction-fetch contention experiment)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/cfg_summary.txt
for cfg in 256x1 256x2 256x3 512x1 512x2 1024x1; do
  VGSIM_TAU_CFG=$cfg VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$cfg.log 2> gpurun_out/bench_$cfg.err
  ms=$(python -c "import json;d=json.loads(open('gpurun_out/bench_$cfg.log').read().strip().splitlines()[-1]);print(d['roofline']['kernel_ms'], d['leaps_per_s'], d['device_error_flags'])" 2>/dev/null || echo fail)
  echo "cfg $cfg kernel_ms/leaps_per_s/err $ms" >> gpurun_out/cfg_summary.txt
done
VGSIM_TAU_CFG=1024x1 timeout 600 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_tau_1024.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau_1024.log
cat gpurun_out/cfg_summary.txt; tail -3 gpurun_out/pytest_tau_1024.log
