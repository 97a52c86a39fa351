#!/bin/bash
# lockstep cadence A/B: the warps of a CTA meet every N-th leap
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_summary.txt
run() {  # name, env...
  name=$1; shift
  env "$@" VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]);print('$name', d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'], d['device_error_flags'])" >> gpurun_out/ab_summary.txt 2>&1 || tail -3 gpurun_out/bench_$name.err >> gpurun_out/ab_summary.txt
}
run every1 VGSIM_TAU_SYNC_EVERY=1
run every2 VGSIM_TAU_SYNC_EVERY=2
run every4 VGSIM_TAU_SYNC_EVERY=4
run every8 VGSIM_TAU_SYNC_EVERY=8
run every2s1 VGSIM_TAU_SYNC_EVERY=2 VGSIM_TAU_SYNC=1
run every4s1 VGSIM_TAU_SYNC_EVERY=4 VGSIM_TAU_SYNC=1
cat gpurun_out/ab_summary.txt
VGSIM_TAU_SYNC_EVERY=4 timeout 300 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 300 -k "warp_kernel or schedule" > gpurun_out/pytest_tau_every4.log 2>&1
tail -3 gpurun_out/pytest_tau_every4.log
