#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 100 python scripts/run_world.py 64 4000 30000 40 101 > $O/r2k_world64_30k.json 2> $O/r2k_world64_30k.err; tail -c 1200 $O/r2k_world64_30k.json; tail -6 $O/r2k_world64_30k.err
timeout 200 python scripts/run_world.py 32 4000 100000 40 101 > $O/r2k_world32_100k.json 2> $O/r2k_world32_100k.err; tail -c 1200 $O/r2k_world32_100k.json; tail -6 $O/r2k_world32_100k.err
timeout 300 python bench.py --no-cpu-baseline --no-curves --steps 5 --warmup 3 --phases > $O/r2k_bench_phases.json 2> $O/r2k_bench_phases.err; tail -3 $O/r2k_bench_phases.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2k_bench_phases.json") if l.startswith("{")][-1])
print(" | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]))
print(j.get("tau_phase_cycles_per_leap"))
for w in j["windows"]: print(w.get("t"), w.get("phase_cycles_per_leap"))
PY
