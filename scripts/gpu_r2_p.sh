#!/bin/bash
# round 2, call p: A/B of tau-kernel variants (loop scalars in smem; 12 warps x 168 regs; 15 warps; 16 warps with a 64-entry queue)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
B="--no-cpu-baseline --no-curves --steps 5 --warmup 3"
for v in base ls w12r168 w15 w16q64; do
VGSIM_B200_LIB=$PWD/vgsim_b200/libvgsim_b200_$v.so timeout 300 python bench.py $B > $O/r2p_bench_$v.json 2> $O/r2p_bench_$v.err || tail -3 $O/r2p_bench_$v.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_bench*.json")):
    try:
        j=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, " | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]), "direct %.0f ms" % (j["direct"]["kernel_ms"]))
    except Exception as e: print(f, "failed", e)
PY
