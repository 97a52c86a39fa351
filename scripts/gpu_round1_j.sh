#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/cfg_summary.txt
timeout 600 python -m pytest tests/test_gpu_tau.py tests/test_gpu_genealogy.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
tail -3 gpurun_out/pytest_tau.log
for cfg in 4x1 2x1 1x1; do
  VGSIM_TAU_CFG=$cfg VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --phases > gpurun_out/bench_$cfg.log 2> gpurun_out/bench_$cfg.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$cfg.log').read().strip().splitlines()[-1]);print('$cfg', d['roofline']['kernel_ms'], d['leaps_per_s'], d['device_error_flags'], json.dumps({k:int(v) for k,v in d['tau_phase_cycles_per_leap'].items()}))" >> gpurun_out/cfg_summary.txt 2>&1 || tail -3 gpurun_out/bench_$cfg.err >> gpurun_out/cfg_summary.txt
done
cat gpurun_out/cfg_summary.txt
