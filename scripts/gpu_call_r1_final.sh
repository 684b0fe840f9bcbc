#!/bin/bash
# The round's last GPU call (10 GPU-minutes left): A/B the scan variants, parity of v4, a bench line with v4,
# the GPU tests that have not run on hardware yet, one ncu capture of v4, then as much of the full suite under v4 as fits.
mkdir -p gpurun_out
exec > >(tee gpurun_out/final_call.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== A/B timing (ps, ph)"; date
timeout 200 python scripts/time_scan_variants.py --model ps,ph | tee gpurun_out/ab_scan.jsonl
echo "== v4 parity + vllm anchor + tests of the last commit"; date
timeout 300 python -m pytest tests/test_gpu_scan_v4.py tests/test_gpu_parity.py tests/test_batch_prep.py tests/test_gpu_shard_train.py \
    -m gpu -q -rs --timeout 120 -k "v4 or vllm or batch_prep or shard_train" 2>&1 | tail -25
echo "== bench, forward, scan variant 4"; date
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant 4 | tee gpurun_out/bench_ps_v4.json
echo "== bench, forward, scan variant 3 (same box)"; date
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant 3 | tee gpurun_out/bench_ps_v3.json
echo "== ncu --set full of the v4 kernel"; date
timeout 240 ncu --set full --clock-control none --import-source on -k regex:bimamba_scan_fwd_v4 -s 2 -c 1 -f -o gpurun_out/scan_v4 \
    python scripts/time_scan_variants.py --model ps --variants 4 --iters 1 > gpurun_out/ncu_v4.log 2>&1
ls -la gpurun_out | tail -5
echo "== full GPU suite with the scan defaulting to variant 4 where it applies"; date
CAD_SCAN_VARIANT=4 timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
date
