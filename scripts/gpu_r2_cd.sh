#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_direct.py tests/test_ts_tables.py tests/test_gpu_example_script.py tests/test_gpu_sweep.py tests/test_cli_io.py -q -m gpu --timeout 500 > $O/r2_cd_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_cd_pytest.log; tail -12 $O/r2_cd_pytest.log | cut -c1-300
