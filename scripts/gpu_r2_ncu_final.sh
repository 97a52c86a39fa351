#!/bin/bash
# round 2: ncu captures of the tau warp kernel on the final build, sparse (t=60) and dense (t=120) windows of the bench trajectories
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for T in 60 120; do
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2_final_prof_tau$T -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60,90,120 --profile-window $T > $O/r2_final_ncu_tau$T.log 2>&1
tail -1 $O/r2_final_ncu_tau$T.log
done
