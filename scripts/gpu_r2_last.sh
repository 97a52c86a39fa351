#!/bin/bash
# round 2, last check of the committed build: full GPU suite + smoke
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/r2_last_pytest.log 2>&1
echo "pytest exit $?" >> $O/r2_last_pytest.log; tail -4 $O/r2_last_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
