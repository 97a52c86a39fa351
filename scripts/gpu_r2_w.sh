#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nproc
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q -m gpu --timeout 1200 --durations=8 > $O/r2w_pytest_scale.log 2>&1
echo "pytest exit $?" >> $O/r2w_pytest_scale.log; tail -25 $O/r2w_pytest_scale.log | cut -c1-300
