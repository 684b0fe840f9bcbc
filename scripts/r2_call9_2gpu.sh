#!/bin/bash
# Round-2 GPU call 9 (2 GPUs): real NVLink peers — symmetric memory rendezvous, peer-exchange kernels between two processes,
# the sequence-sharded bench (graph + peer exchange; NCCL exchange; replicas; 16k shards on the time-parallel kernel), sharded train step.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call9.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader; nvidia-smi topo -m | head -5
echo "== multi-rank tests"; date
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -12
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$1 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline "${@:2}" 2>gpurun_out/r2c9_err_$1.log | grep '^{' ; tail -3 gpurun_out/r2c9_err_$1.log | cut -c1-300; }
echo "== bench N=2 sharded, peer exchange + graph (default)"; date
run 1 | tee gpurun_out/r2c9_bench_n2_seq_peer_graph.json
echo "== bench N=2 sharded, peer exchange, eager"; date
run 2 --no-graph | tee gpurun_out/r2c9_bench_n2_seq_peer_eager.json
echo "== bench N=2 sharded, NCCL exchange, eager"; date
run 3 --exchange nccl | tee gpurun_out/r2c9_bench_n2_seq_nccl.json
echo "== bench N=2 sharded, 16k-token shards (the per-rank shape of the 8-way split: time-parallel kernel)"; date
run 4 --seqlen 32768 | tee gpurun_out/r2c9_bench_n2_seq_L32k.json
echo "== bench N=2 replicas"; date
run 5 --shard none | tee gpurun_out/r2c9_bench_n2_replicas.json
echo "== bench N=2 sharded Ph train step (configs[3])"; date
run 6 --model ph --mode train --steps 5 | tee gpurun_out/r2c9_bench_n2_train_ph.json
echo "== bench N=1 for the same box"; date
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2c9_bench_n1.json
date
