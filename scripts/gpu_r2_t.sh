#!/bin/bash
# round 2, call t (8 GPUs): BASELINE configs[4] -- the 65,536-point R0 x migration sweep over 8 ranks with the timed NCCL all-gather
# of the summaries -- and the bench at N = 8
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/run_sweep.py 20000 64 > $O/r2t_sweep8.json 2> $O/r2t_sweep8.err
echo "sweep exit $?"; tail -c 1500 $O/r2t_sweep8.json; tail -3 $O/r2t_sweep8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-curves > $O/r2t_bench_n8.json 2> $O/r2t_bench_n8.err
echo "bench exit $?"; tail -c 600 $O/r2t_bench_n8.json; tail -3 $O/r2t_bench_n8.err
