#!/bin/bash
# round 2, call v: full GPU suite + smoke + world config after the decode / curves changes
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/r2v_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2v_pytest.log; tail -6 $O/r2v_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python scripts/run_world.py 256 4000 100000 40 101 > $O/r2v_world256_100k.json 2> $O/r2v_world256_100k.err; tail -c 1300 $O/r2v_world256_100k.json; tail -6 $O/r2v_world256_100k.err
