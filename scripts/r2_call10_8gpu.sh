#!/bin/bash
# Round-2 GPU call 10 (8 GPUs): the strong-scaling curve of ONE 131k-token Caduceus-PS forward, sharded on the sequence axis
# (peer-memory exchange + CUDA graph), N = 8, 4, 2, 1 on one box; NCCL-exchange line at 8 for comparison.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call10.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader | sort | uniq -c
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/r2c10_err.log | grep '^{' ; tail -2 gpurun_out/r2c10_err.log | cut -c1-300; }
echo "== N=8 sharded (default: peer exchange + graph)"; date
run 8 | tee gpurun_out/r2c10_bench_n8.json
echo "== N=4"; date
run 4 | tee gpurun_out/r2c10_bench_n4.json
echo "== N=2"; date
run 2 | tee gpurun_out/r2c10_bench_n2.json
echo "== N=1"; date
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2c10_bench_n1.json
echo "== N=8 sharded, NCCL exchange (eager)"; date
run 8 --exchange nccl | tee gpurun_out/r2c10_bench_n8_nccl.json
echo "== N=8 Ph sharded"; date
run 8 --model ph | tee gpurun_out/r2c10_bench_n8_ph.json
date
