#!/bin/bash
# Round-2 GPU call 12 (2 GPUs): segment-parallel carry fix-up of a shard scanned by the time-parallel kernel (the 8-way split's
# per-rank shape): parity, then the 16k-shard bench line against call 9's 10.22 ms, the default N=2 line, and per-kernel pieces.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call12.log) 2>&1
date
echo "== parity of the new fix-up flow + head/CE + multi-rank"; date
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "fixup or sharding" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q --timeout 300 -k "fused" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 400 -k "real_peers or collective" 2>&1 | tail -3
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2970$1 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline "${@:2}" 2>gpurun_out/r2c12_err_$1.log | grep '^{' ; tail -2 gpurun_out/r2c12_err_$1.log | cut -c1-300; }
echo "== bench N=2, 16k-token shards"; date
run 1 --seqlen 32768 | tee gpurun_out/r2c12_bench_n2_seq_L32k.json
echo "== bench N=2 default"; date
run 2 | tee gpurun_out/r2c12_bench_n2.json
echo "== per-kernel pieces of one rank at the shard lengths (GPU 0)"; date
timeout 200 python scripts/time_shard_pieces.py | tee gpurun_out/r2c12_shard_pieces.jsonl
date
