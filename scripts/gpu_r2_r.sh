#!/bin/bash
# round 2, call r: new tau-tolerance / curves tests, smoke(), and A/B of the aggregation thresholds and queue sizes
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_tau.py tests/test_gpu_curves.py -q -m gpu --timeout 500 > $O/r2r_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2r_pytest.log; tail -6 $O/r2r_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
for v in thmig05 thmut05 th015 th04 q80 q112; do
VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 300 python bench.py $B > $O/r2r_bench_$v.json 2> $O/r2r_bench_$v.err || tail -3 $O/r2r_bench_$v.err
done
timeout 300 python bench.py $B > $O/r2r_bench_base.json 2> $O/r2r_bench_base.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]))
    except Exception as e: print(f, "failed", e)
PY
