#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_direct.py tests/test_gpu_genealogy.py -q -m gpu --timeout 500 > $O/r2x_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2x_pytest.log; tail -25 $O/r2x_pytest.log | cut -c1-400
