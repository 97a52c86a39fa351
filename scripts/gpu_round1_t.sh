#!/bin/bash
# generic env A/B: each argument is "name:VAR=val,VAR=val"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_summary.txt
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $(echo $envs | tr ',' ' ') VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]);print('$name', d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'], d['device_error_flags'], d['value'], d.get('tau_phase_cycles_per_leap'))" >> gpurun_out/ab_summary.txt 2>&1 || tail -3 gpurun_out/bench_$name.err >> gpurun_out/ab_summary.txt
done
cat gpurun_out/ab_summary.txt
