#!/bin/bash
# warp-per-replicate tau kernel: parity (incl. warp vs team), bench with the phase tap, warps-per-CTA A/B, one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tau.py -q -m gpu --timeout 600 > gpurun_out/pytest_tau.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tau.log
tail -25 gpurun_out/pytest_tau.log
for w in 16 14 12 8; do
  VGSIM_TAU_WARPS=$w VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_w$w.json').read().strip().splitlines()[-1]);print('warps $w', d['roofline']['kernel_ms'], d['leaps_per_s'], d['roofline']['frac'], d['device_error_flags'], d['events_per_leap'])" || tail -3 gpurun_out/bench_w$w.err
done
VGSIM_BENCH_WATCHDOG=100 timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --phases > gpurun_out/bench_phases.json 2> gpurun_out/bench_phases.err
python -c "import json;d=json.loads(open('gpurun_out/bench_phases.json').read().strip().splitlines()[-1]);print(d['roofline']['kernel_ms'], d['tau_phase_cycles_per_leap'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tau_warp_kernel -s 1 -c 1 -f -o gpurun_out/prof_tauw \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
