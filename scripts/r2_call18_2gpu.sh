#!/bin/bash
# Round-2 GPU call 18 (2 GPUs): the sequence-sharded forward (peer-memory exchange + CUDA graph) with the tcgen05 projection kernel
# inside the captured graph: default N=2 line and the 16k-token-shard shape (the per-rank shape of the 8-way split).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call18.log) 2>&1
date
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2971$1 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline "${@:2}" 2>gpurun_out/r2c18_err_$1.log | grep '^{' ; tail -2 gpurun_out/r2c18_err_$1.log | cut -c1-300; }
echo "== bench N=2 default"; date
run 1 | tee gpurun_out/r2c18_bench_n2.json
echo "== bench N=2, 16k-token shards"; date
run 2 --seqlen 32768 | tee gpurun_out/r2c18_bench_n2_seq_L32k.json
echo "== real peers over NVLink: sharded forward == unsharded"; date
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 -k "real_peers" 2>&1 | tail -3
date
