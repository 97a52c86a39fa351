#!/bin/bash
# single sampler call site + constant tables: bench, then tau parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/ab_summary.txt
run() {  # name, env...
  name=$1; shift
  env "$@" VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]);print('$name', d['roofline']['kernel_ms'], round(d['roofline']['frac'],4), d['ms_per_step'], d['device_error_flags'], d['value'])" >> gpurun_out/ab_summary.txt 2>&1 || tail -3 gpurun_out/bench_$name.err >> gpurun_out/ab_summary.txt
}
run base A=1
for v in $VARIANTS; do run $v VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so; done
cat gpurun_out/ab_summary.txt
if [ -n "$TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
tail -5 gpurun_out/pytest_tau.log
fi
