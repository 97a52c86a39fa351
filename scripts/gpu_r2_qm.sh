#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
for v in q112m05 q112; do
VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 300 python bench.py $B > $O/r2_qm_bench_$v.json 2> $O/r2_qm_bench_$v.err || tail -3 $O/r2_qm_bench_$v.err
done
timeout 300 python bench.py $B > $O/r2_qm_bench_base.json 2> $O/r2_qm_bench_base.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_qm_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]))
    except Exception as e: print(f, "failed", e)
PY
