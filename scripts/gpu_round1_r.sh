#!/bin/bash
# A/B: lockstep group size, 144 registers (__maxnreg__), sampler tables in constant memory
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_summary.txt
run() {  # name, env...
  name=$1; shift
  env "$@" VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]);print('$name', d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'], d['device_error_flags'], d['value'])" >> gpurun_out/ab_summary.txt 2>&1 || tail -3 gpurun_out/bench_$name.err >> gpurun_out/ab_summary.txt
}
run base A=1
run g7 VGSIM_TAU_GROUP=7
run g4 VGSIM_TAU_GROUP=4
run g2 VGSIM_TAU_GROUP=2
run r144 VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_r144.so
run ct VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_ct.so
run r144ct VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_r144ct.so
run r144ctg7 VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_r144ct.so VGSIM_TAU_GROUP=7
cat gpurun_out/ab_summary.txt
