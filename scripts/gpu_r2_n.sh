#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_genealogy.py tests/test_gpu_archive.py -q -m gpu --timeout 300 > $O/r2n_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2n_pytest.log; tail -4 $O/r2n_pytest.log
timeout 100 python scripts/run_world.py 32 4000 100000 40 101 > $O/r2n_world32_100k.json 2> $O/r2n_world32_100k.err; tail -c 900 $O/r2n_world32_100k.json; tail -3 $O/r2n_world32_100k.err
timeout 240 python scripts/run_world.py 256 4000 100000 40 101 > $O/r2n_world256_100k.json 2> $O/r2n_world256_100k.err; tail -c 2500 $O/r2n_world256_100k.json; tail -7 $O/r2n_world256_100k.err
