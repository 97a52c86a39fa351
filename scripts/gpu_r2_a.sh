#!/bin/bash
# round 2, call a: baseline with the windowed bench, full GPU suite, ncu of direct / genealogy / tau at the dense window
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > $O/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/r2a_pytest_gpu.log
tail -5 $O/r2a_pytest_gpu.log
timeout 900 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
tail -c 6000 $O/r2a_bench.json; tail -5 $O/r2a_bench.err
timeout 300 python scripts/profile_dg.py 2048 55 > $O/r2a_dg.log 2>&1; cat $O/r2a_dg.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 -o $O/r2a_prof_direct -f python scripts/profile_dg.py 1024 50 > $O/r2a_ncu_direct.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:genealogy_kernel -c 1 -o $O/r2a_prof_gen -f python scripts/profile_dg.py 1024 50 > $O/r2a_ncu_gen.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2a_prof_tau120 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60,120 --profile-window 120 > $O/r2a_ncu_tau120.log 2>&1
ls -la $O/*.ncu-rep | tail -5
