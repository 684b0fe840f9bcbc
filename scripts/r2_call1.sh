#!/bin/bash
# Round-2 GPU call 1: lane = channel scan (variants 20-23) on hardware for the first time, the gated variant tests,
# the upstream sm_100 kernel timed next to ours, one ncu --set full capture of v20.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call1.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== hw_probe A (fwd variants) + D (v20 pipeline)"; date
timeout 120 ./scripts/_bin/hw_probe 131072 AD
cp gpurun_out/hw_probe.log gpurun_out/r2_hw_probe_AD.log
echo "== gated variant tests on hardware"; date
CAD_RUN_UNMEASURED=1 timeout 400 python -m pytest tests/test_gpu_scan_variants.py -m gpu -q --timeout 120 -x 2>&1 | tail -15
echo "== upstream kernel (vLLM sm_100 build) timing"; date
timeout 200 python scripts/time_vllm_scan.py | tee gpurun_out/r2_vllm_scan.jsonl
echo "== ours, same box"; date
timeout 200 python scripts/time_scan_variants.py --model ps,ph --variants 3,11,12,20,21,22,23 | tee gpurun_out/r2_ab_scan.jsonl
echo "== ncu --set full of v20 through the probe"; date
timeout 240 ncu --set full --clock-control none --import-source on --target-processes all -k regex:bimamba_scan_fwd_v20 -s 2 -c 1 -f \
    -o gpurun_out/r2_scan_v20 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2_ncu_v20.log 2>&1
tail -5 gpurun_out/r2_ncu_v20.log
echo "== bench under 20 / 22"; date
for v in 20 22; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant $v | tee gpurun_out/r2_bench_ps_scan_v$v.json
done
date
