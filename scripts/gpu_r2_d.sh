#!/bin/bash
# round 2, call d: new direct kernel + choose tap, drift enumeration, parked splits / two-pass PTRS, pipeline fixes
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x --deselect tests/test_gpu_parity_scale.py > $O/r2d_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/r2d_pytest_gpu.log
tail -12 $O/r2d_pytest_gpu.log
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 600 python bench.py $B > $O/r2d_bench.json 2> $O/r2d_bench.err; tail -3 $O/r2d_bench.err
for v in thg1 thg2 thg4 thg8; do
  VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 600 python bench.py $B > $O/r2d_bench_$v.json 2> $O/r2d_bench_$v.err; tail -2 $O/r2d_bench_$v.err
done
for sy in 0 1 7; do
  VGSIM_TAU_SYNC=$sy VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_dbg.so timeout 600 python bench.py $B > $O/r2d_bench_sync$sy.json 2> $O/r2d_bench_sync$sy.err; tail -2 $O/r2d_bench_sync$sy.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2d_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "ms/step %.2f e2e %.2f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]), "direct %.0f ms %.3g ev/s" % (j["direct"]["kernel_ms"], j["direct"]["events_per_s"]))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python -m pytest tests/test_gpu_parity_scale.py -q -m gpu --timeout 900 -k "lockdown or full_length" > $O/r2d_pytest_parity.log 2>&1; tail -5 $O/r2d_pytest_parity.log
