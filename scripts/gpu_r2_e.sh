#!/bin/bash
# round 2, call e: A/B of the three tau changes of call d + wipe pacing + theta 0.1; ncu of tau (t=60) and of the new direct kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 600 python bench.py $B > $O/r2e_bench.json 2> $O/r2e_bench.err; tail -3 $O/r2e_bench.err
for v in b0 park0 ptrs0 all0 wipe3 th01; do
  VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 600 python bench.py $B > $O/r2e_bench_$v.json 2> $O/r2e_bench_$v.err; tail -2 $O/r2e_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2e_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "ms/step %.2f e2e %.2f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]), "direct %.0f ms" % (j["direct"]["kernel_ms"]))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2e_prof_tau60 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60 --profile-window 60 > $O/r2e_ncu_tau60.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 -o $O/r2e_prof_direct -f python scripts/profile_dg.py 1024 50 > $O/r2e_ncu_direct.log 2>&1
tail -2 $O/r2e_ncu_direct.log
