#!/bin/bash
# full GPU parity suite, default bench (both arms), phase tap, ncu launch list, one ncu --set full capture of tau_kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
VGSIM_BENCH_WATCHDOG=200 timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
VGSIM_BENCH_WATCHDOG=200 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --phases > gpurun_out/bench_phases.json 2> gpurun_out/bench_phases.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_kernel -s 1 -c 1 -f -o gpurun_out/prof_tau \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out
