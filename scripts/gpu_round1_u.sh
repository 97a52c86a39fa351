#!/bin/bash
# full GPU suite + default bench (with cpu baseline) + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 2500 gpurun_out/bench.json
if [ -n "$REFARM" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1200 gpurun_out/bench_reference.json
fi
