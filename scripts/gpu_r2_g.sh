#!/bin/bash
# round 2, call g: two-half queue push (qcap 96 default vs 128 / 64), lane-batched RNG in the direct kernel, tests with skip reasons
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tau.py tests/test_gpu_direct.py tests/test_gpu_choose.py tests/test_writers_parity.py tests/test_gpu_example_script.py tests/test_gpu_genealogy.py -q -m gpu -rs --timeout 600 > $O/r2g_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2g_pytest.log
tail -12 $O/r2g_pytest.log
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 600 python bench.py $B > $O/r2g_bench.json 2> $O/r2g_bench.err; tail -3 $O/r2g_bench.err
for v in q128 q64; do
VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 600 python bench.py $B > $O/r2g_bench_$v.json 2> $O/r2g_bench_$v.err
done
VGSIM_TAU_SYNC=7 VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_dbg.so timeout 600 python bench.py $B > $O/r2g_bench_sync7.json 2> $O/r2g_bench_sync7.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "ms/step %.2f e2e %.2f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]), "direct %.0f ms %.3g ev/s" % (j["direct"]["kernel_ms"], j["direct"]["events_per_s"]))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tau_warp_kernel -c 1 -o $O/r2g_prof_tau60 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-curves --windows 60 --profile-window 60 > $O/r2g_ncu_tau60.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:direct_kernel -c 1 -o $O/r2g_prof_direct -f python scripts/profile_dg.py 1024 50 > $O/r2g_ncu_direct.log 2>&1
grep direct: $O/r2g_ncu_direct.log
