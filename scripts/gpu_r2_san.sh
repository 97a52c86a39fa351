#!/bin/bash
# round 2: compute-sanitizer memcheck over the new code paths (archive kernels, archived-leap readers, blocked tau call, clamped replay)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_archive.py -q -m gpu --timeout 800 -k "s9 or reset or (blocks and t3small and warp)" > $O/r2_san_archive.log 2>&1
echo "sanitizer exit $?" >> $O/r2_san_archive.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|exit" $O/r2_san_archive.log | tail -8
