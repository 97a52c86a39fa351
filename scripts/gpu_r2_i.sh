#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_archive.py -q -m gpu --timeout 200 > $O/r2i_pytest_archive.log 2>&1
echo "pytest exit $?" >> $O/r2i_pytest_archive.log
tail -3 $O/r2i_pytest_archive.log | cut -c1-400
timeout 120 python scripts/run_world.py 32 1200 100000 40 101 > $O/r2i_world32_blocks.json 2> $O/r2i_world32_blocks.err; tail -c 1500 $O/r2i_world32_blocks.json; tail -3 $O/r2i_world32_blocks.err
timeout 100 python scripts/run_world.py 256 600 100000 40 101 > $O/r2i_world256_600.json 2> $O/r2i_world256_600.err; tail -c 1500 $O/r2i_world256_600.json; tail -5 $O/r2i_world256_600.err
