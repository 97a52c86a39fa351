#!/bin/bash
# round 2, call c: async e2e, theta / queue / warps variants of the new warp kernel, ncu at t=60 and t=120, new parity tests
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 600 python bench.py $B > $O/r2c_bench.json 2> $O/r2c_bench.err; tail -2 $O/r2c_bench.err
for v in th01 th05 q96 q192 w15 w13; do
  VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 600 python bench.py $B > $O/r2c_bench_$v.json 2> $O/r2c_bench_$v.err; tail -2 $O/r2c_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "ms/step %.2f e2e %.2f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2c_prof_tau60 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60 --profile-window 60 > $O/r2c_ncu_tau60.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2c_prof_tau120 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60,120 --profile-window 120 > $O/r2c_ncu_tau120.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity_scale.py -q -m gpu --timeout 900 > $O/r2c_pytest_parity.log 2>&1
echo "pytest exit $?" >> $O/r2c_pytest_parity.log
tail -15 $O/r2c_pytest_parity.log
