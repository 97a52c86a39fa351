#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_tau.py -x -q -m gpu > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
timeout 600 python scripts/probe_tau.py t3 4096 32 70 > gpurun_out/probe_t3.log 2>&1
tail -5 gpurun_out/pytest_tau.log; cat gpurun_out/probe_t3.log
