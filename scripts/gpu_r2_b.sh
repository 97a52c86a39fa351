#!/bin/bash
# round 2, call b: canonical draw/drain of the warp kernel -- parity with the team kernel, windows, theta / queue variants
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 600 -x > $O/r2b_pytest_tau.log 2>&1
echo "pytest exit $?" >> $O/r2b_pytest_tau.log
tail -4 $O/r2b_pytest_tau.log
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 600 python bench.py $B > $O/r2b_bench.json 2> $O/r2b_bench.err; tail -2 $O/r2b_bench.err
for v in thg1 thg3 thg8 q256; do
  VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 600 python bench.py $B > $O/r2b_bench_$v.json 2> $O/r2b_bench_$v.err; tail -2 $O/r2b_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]))
    except Exception as e: print(f, "failed", e)
PY
