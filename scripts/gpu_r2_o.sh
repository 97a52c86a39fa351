#!/bin/bash
# round 2, call o: ncu captures of the tau warp kernel at the sparse (t=60) and dense (t=120) windows, current build
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for T in 60 120; do
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2o_prof_tau$T -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows $T --profile-window $T > $O/r2o_ncu_tau$T.log 2>&1
tail -2 $O/r2o_ncu_tau$T.log
done
ls -la $O/r2o_prof_tau*
