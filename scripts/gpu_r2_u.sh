#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_archive.py -q -m gpu --timeout 200 > $O/r2u_pytest_archive.log 2>&1
echo "pytest exit $?" >> $O/r2u_pytest_archive.log; tail -3 $O/r2u_pytest_archive.log | cut -c1-300
timeout 240 python scripts/run_world.py 256 4000 100000 40 101 > $O/r2u_world256_100k.json 2> $O/r2u_world256_100k.err; tail -c 1600 $O/r2u_world256_100k.json; tail -6 $O/r2u_world256_100k.err
