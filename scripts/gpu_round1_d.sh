#!/bin/bash
# hang hunt: v3 kernel first (current build), then the committed v2 build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/debug_hang.py > gpurun_out/dbg_v3.log 2>&1
echo "exit $?" >> gpurun_out/dbg_v3.log
timeout 700 python scripts/debug_hang.py --lib gpurun_dbg_libv2.so > gpurun_out/dbg_v2.log 2>&1
echo "exit $?" >> gpurun_out/dbg_v2.log
timeout 600 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
cat gpurun_out/dbg_v3.log; cat gpurun_out/dbg_v2.log; tail -15 gpurun_out/pytest_tau.log
