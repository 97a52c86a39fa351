#!/bin/bash
# round 2, call y: balanced (LPT) assignment of replicate groups to CTAs vs the boustrophedon pairing
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 200 -k "schedule or reproduces" > $O/r2y_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2y_pytest.log; tail -4 $O/r2y_pytest.log | cut -c1-300
echo "--- tail, balanced"; timeout 100 python scripts/probe_tail.py 4096 60 0 | grep -v "busiest\|within"
echo "--- tail, boustrophedon"; timeout 100 python scripts/probe_tail.py 4096 60 64 | grep -v "busiest\|within"
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 300 python bench.py $B > $O/r2y_bench.json 2> $O/r2y_bench.err || tail -3 $O/r2y_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2y_bench.json") if l.startswith("{")][-1])
print(" | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "launches", j["gpu_launches"])
PY
