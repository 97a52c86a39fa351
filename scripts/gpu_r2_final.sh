#!/bin/bash
# round 2, closing run on the final build: full GPU suite, smoke, bench (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/r2_final_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_final_pytest.log; tail -5 $O/r2_final_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -2 $O/r2_final_bench.err
timeout 600 python bench.py --impl reference > $O/r2_final_bench_reference.json 2> $O/r2_final_bench_reference.err; tail -2 $O/r2_final_bench_reference.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_final_bench.json") if l.startswith("{")][-1])
print("value %.4g e2e %.4g frac %.4f" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"]), " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "direct %.0f ms" % j["direct"]["kernel_ms"], "curves", j["epidemic_curves"].get("achieved_GBps"))
r=json.loads([l for l in open("gpurun_out/r2_final_bench_reference.json") if l.startswith("{")][-1])
print("reference %.4g" % r["value"], "ratio e2e %.0f" % (j["e2e"]["value"]/r["e2e"]["value"]))
PY
