#!/bin/bash
# round 2, call h: sparse archive (tests), world config at its stated size in leap blocks, full GPU suite, bench
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_archive.py -q -m gpu -x --timeout 500 > $O/r2h_pytest_archive.log 2>&1
echo "pytest exit $?" >> $O/r2h_pytest_archive.log
tail -15 $O/r2h_pytest_archive.log
timeout 300 python scripts/run_world.py 32 1200 100000 40 0 > $O/r2h_world32_dense.json 2> $O/r2h_world32_dense.err; tail -c 1500 $O/r2h_world32_dense.json; tail -3 $O/r2h_world32_dense.err
timeout 300 python scripts/run_world.py 32 1200 100000 40 101 > $O/r2h_world32_blocks.json 2> $O/r2h_world32_blocks.err; tail -c 1500 $O/r2h_world32_blocks.json; tail -3 $O/r2h_world32_blocks.err
timeout 900 python scripts/run_world.py 256 4000 100000 40 101 > $O/r2h_world256.json 2> $O/r2h_world256.err; tail -c 2000 $O/r2h_world256.json; tail -5 $O/r2h_world256.err
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 --deselect tests/test_gpu_archive.py > $O/r2h_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2h_pytest.log
tail -8 $O/r2h_pytest.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > $O/r2h_bench.json 2> $O/r2h_bench.err; tail -3 $O/r2h_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r2h_bench.json") if l.startswith("{")][-1])
print(" | ".join("t=%g %.2f ms frac %.3f" % (w["t"], w["kernel_ms"], w["frac"]) for w in j["windows"]), "e2e/value %.3f" % (j["e2e"]["value"]/j["value"]))
PY
