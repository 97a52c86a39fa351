#!/bin/bash
# warp kernel v3 (pipelined row wipe, 144 registers, sampler fast path): parity, A/B, ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_summary.txt
timeout 900 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 600 > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
tail -12 gpurun_out/pytest_tau.log
run() {  # name, env...
  name=$1; shift
  env "$@" VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --phases > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]);print('$name', d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'], d['device_error_flags'], {k:int(v) for k,v in d['tau_phase_cycles_per_leap'].items() if v})" >> gpurun_out/ab_summary.txt 2>&1 || tail -3 gpurun_out/bench_$name.err >> gpurun_out/ab_summary.txt
}
run default A=1
run sync7 VGSIM_TAU_SYNC=7
run sync1 VGSIM_TAU_SYNC=1
cat gpurun_out/ab_summary.txt
VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_tauw5 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out | tail -3
