#!/bin/bash
# round-1 closing run: full GPU suite, default bench (both arms), launch list and one full ncu capture of the tau kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 900 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_tauw_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves > gpurun_out/bench_ncu_full_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:curves_kernel -c 1 -f -o gpurun_out/prof_curves_final \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_curves_final.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
