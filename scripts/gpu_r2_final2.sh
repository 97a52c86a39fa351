#!/bin/bash
# round 2, closing bench on the final build (Mg table in), with the launch list
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py > $O/r2_final2_bench.json 2> $O/r2_final2_bench.err; tail -2 $O/r2_final2_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_final2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_final2_launches.log 2>&1
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_final2_bench.json") if l.startswith("{")][-1])
print("value %.4g e2e %.4g frac %.4f" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"]), " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "direct %.0f ms" % j["direct"]["kernel_ms"], "curves", j["epidemic_curves"].get("achieved_GBps"), "cpu", j["cpu_baseline"]["value"])
PY
