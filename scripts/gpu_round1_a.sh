#!/bin/bash
# first GPU pass of the session: parity suite, smoke, bench (both arms), launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --replicates 1024 --leaps 8 > gpurun_out/bench_ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_ref.log | tail -2
