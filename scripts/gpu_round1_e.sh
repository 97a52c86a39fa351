#!/bin/bash
# v3 tau kernel: parity suite, bench under a watchdog, launch list, ncu --set full on a one-wave launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
VGSIM_BENCH_WATCHDOG=150 timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_kernel -s 1 -c 1 -o gpurun_out/prof_tau \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --replicates 2368 --leaps 16 > gpurun_out/bench_ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
