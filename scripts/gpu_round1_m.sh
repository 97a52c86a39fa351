#!/bin/bash
# warp kernel: parity, lockstep (VGSIM_TAU_SYNC) A/B with the phase tap, ncu capture of the fastest setting
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 600 > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
tail -12 gpurun_out/pytest_tau.log
VGSIM_TAU_SYNC=7 timeout 600 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 300 -k "warp_kernel or agree" > gpurun_out/pytest_tau_sync7.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau_sync7.log
tail -4 gpurun_out/pytest_tau_sync7.log
best=0; bestms=1e9
for sy in 0 1 3 5 7; do
  VGSIM_TAU_SYNC=$sy VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --phases > gpurun_out/bench_s$sy.json 2> gpurun_out/bench_s$sy.err
  ms=$(python -c "import json;d=json.loads(open('gpurun_out/bench_s$sy.json').read().strip().splitlines()[-1]);print(d['roofline']['kernel_ms']);import sys;sys.stderr.write('sync $sy %s %s %s %s\n'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['device_error_flags'], {k:int(v) for k,v in d['tau_phase_cycles_per_leap'].items()}))" 2>> gpurun_out/sync_summary.txt || echo 1e9)
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)"; then best=$sy; bestms=$ms; fi
done
echo "best $best $bestms" >> gpurun_out/sync_summary.txt
cat gpurun_out/sync_summary.txt
VGSIM_TAU_SYNC=$best VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err
VGSIM_TAU_SYNC=$best timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_tauw2 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out | tail -4
