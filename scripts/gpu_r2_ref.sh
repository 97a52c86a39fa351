#!/bin/bash
# the reference arm exactly as the driver launches it at N = 2
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nproc
S=$(date +%s)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $O/r2_ref_n2.json 2> $O/r2_ref_n2.err
echo "exit $? after $(( $(date +%s) - S )) s; json lines: $(grep -c '^{' $O/r2_ref_n2.json)"
python - <<'PY'
import json
r=json.loads([l for l in open("gpurun_out/r2_ref_n2.json") if l.startswith("{")][-1]); print("ref", r["value"], r.get("impl"), r["n_gpus"], r["cpu_baseline"]["sample"])
PY
