#!/bin/bash
# round 2: the bench at N = 2 exactly as the driver launches it (default steps / warmup), both arms
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2_n2_bench.json 2> $O/r2_n2_bench.err
echo "exit $?"; tail -c 400 $O/r2_n2_bench.json | head -c 400; echo; grep -c "^{" $O/r2_n2_bench.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $O/r2_n2_bench_reference.json 2> $O/r2_n2_bench_reference.err
echo "exit $?"; grep -c "^{" $O/r2_n2_bench_reference.json
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2_n2_bench.json") if l.startswith("{")][-1])
print("N=2 value %.4g e2e %.4g frac %.3f n_gpus %d cpu_baseline %s" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["n_gpus"], "cpu_baseline" in j))
r=json.loads([l for l in open("gpurun_out/r2_n2_bench_reference.json") if l.startswith("{")][-1]); print("ref", r["value"], r.get("impl"))
PY
