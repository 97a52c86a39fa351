#!/bin/bash
# per-phase critical path of the tau kernel with 1 and 3 resident CTAs per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in 256x1 256x3; do
  VGSIM_TAU_CFG=$cfg VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --phases > gpurun_out/phases_$cfg.log 2> gpurun_out/phases_$cfg.err
  python -c "import json;d=json.loads(open('gpurun_out/phases_$cfg.log').read().strip().splitlines()[-1]);print('$cfg', d['roofline']['kernel_ms'], json.dumps(d['tau_phase_cycles_per_leap']))"
done
