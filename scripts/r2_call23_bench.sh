#!/bin/bash
# Round-2 GPU call 23 (2 GPUs): bench lines of the end state — N=1 (D2H of the logits overlapped with the next step in the e2e leg,
# second_kernel entry) on GPU 0, then the sequence-sharded N=2 default.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call23.log) 2>&1
date
echo "== bench N=1"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2c23_err_0.log | grep '^{' | tee gpurun_out/r2c23_bench_n1.json
tail -2 gpurun_out/r2c23_err_0.log | cut -c1-300
echo "== bench N=2 default"; date
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2c23_err_1.log | grep '^{' | tee gpurun_out/r2c23_bench_n2.json
tail -2 gpurun_out/r2c23_err_1.log | cut -c1-300
date
