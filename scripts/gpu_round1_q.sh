#!/bin/bash
# current warp kernel: bench, launch list of the bench command, one full ncu capture of the tau kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
tail -1 gpurun_out/bench_q.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_q.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_q.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_tauw6 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_full_q.log 2>&1
ls -la gpurun_out | tail -4
