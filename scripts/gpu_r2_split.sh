#!/bin/bash
# round 2: two-stage split of aggregated out-migration totals -- parity of the tau path, then the bench windows
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tau.py tests/test_gpu_parity_scale.py -q -m gpu --timeout 800 > $O/r2_split_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_split_pytest.log; tail -5 $O/r2_split_pytest.log | cut -c1-300
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
timeout 300 python bench.py $B > $O/r2_split_bench.json 2> $O/r2_split_bench.err || tail -3 $O/r2_split_bench.err
timeout 300 python bench.py $B > $O/r2_split_bench2.json 2> $O/r2_split_bench2.err || tail -3 $O/r2_split_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_split_bench.json","gpurun_out/r2_split_bench2.json"):
    j=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(" | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]))
PY
