#!/bin/bash
# round 2: queue of 112 entries + out-migration aggregation threshold 0.5 as defaults -- tau parity, full bench, launch list
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tau.py tests/test_gpu_parity_scale.py tests/test_gpu_archive.py tests/test_gpu_sweep.py -q -m gpu --timeout 800 > $O/r2_fin3_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_fin3_pytest.log; tail -3 $O/r2_fin3_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > $O/r2_fin3_bench.json 2> $O/r2_fin3_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_fin3_bench.json") if l.startswith("{")][-1])
print("value %.4g e2e %.4g frac %.4f" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"]), " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "direct %.0f ms" % j["direct"]["kernel_ms"], "curves", j["epidemic_curves"].get("achieved_GBps"))
PY
