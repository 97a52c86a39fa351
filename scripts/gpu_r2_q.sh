#!/bin/bash
# round 2, call q: evidence for profiles/ -- bench (both arms), ncu launch list of the bench command, ncu captures of the
# genealogy kernel (world shape, archived log) and of the tau kernel at the dense window of the bench trajectories
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/r2q_bench.json 2> $O/r2q_bench.err; tail -2 $O/r2q_bench.err
timeout 600 python bench.py --impl reference > $O/r2q_bench_reference.json 2> $O/r2q_bench_reference.err; tail -2 $O/r2q_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2q_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2q_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:genealogy_kernel -c 1 -o $O/r2q_prof_genealogy -f python scripts/run_world.py 32 1200 20000 40 101 > $O/r2q_ncu_genealogy.log 2>&1; tail -2 $O/r2q_ncu_genealogy.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2q_prof_tau120 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60,90,120 --profile-window 120 > $O/r2q_ncu_tau120.log 2>&1; tail -2 $O/r2q_ncu_tau120.log
python - <<'PY'
import json
for f in ("gpurun_out/r2q_bench.json","gpurun_out/r2q_bench_reference.json"):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, j.get("value"), j.get("e2e"), j.get("roofline",{}).get("frac"), j.get("cpu_baseline"))
    except Exception as e: print(f,"failed",e)
PY
