#!/bin/bash
# First GPU call of the next round: pytest-level hardware parity of the kernels whose C-ABI probe (scripts/hw_probe.cu) was green at the end of round 1,
# A/B timings, bench lines under variants 11 / 12 / 3, one ncu capture, the full suite under variant 11.  ~6 minutes on one B200.   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/r2_first_gpu_call.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_first_call.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== Python-free probe: every scan variant vs variant 3 + launch times (4 s)"; date
./scripts/_bin/hw_probe 131072
echo "== hardware parity of scan variants 9..12 (barrier-free hand-over; 16-bit / fp32 tile) and of conv_xproj's bc16 output"; date
CAD_RUN_UNMEASURED=1 timeout 300 python -m pytest tests/test_gpu_scan_variants.py -m gpu -q --timeout 120 2>&1 | tail -8
echo "== A/B timing on the headline shapes"; date
timeout 200 python scripts/time_scan_variants.py --model ps,ph --variants 3,7,9,10,11,12,4 | tee gpurun_out/r2_ab_scan.jsonl
echo "== bench with the scan forced to 11 / 12 / default"; date
for v in 11 12 3; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant $v | tee gpurun_out/r2_bench_ps_scan_v$v.json
done
echo "== ncu --set full of variant 11"; date
timeout 240 ncu --set full --clock-control none --import-source on -k regex:bimamba_scan_fwd_v11 -s 2 -c 1 -f -o gpurun_out/r2_scan_v11 \
    python scripts/time_scan_variants.py --model ps --variants 11 --iters 1 > gpurun_out/r2_ncu_v11.log 2>&1
echo "== ncu --set full of the lane = channel kernel (variant 20, pass A at 37 segments) through the Python-free probe"; date
timeout 240 ncu --set full --clock-control none --import-source on --target-processes all -k regex:bimamba_scan_fwd_v20 -s 2 -c 1 -f \
    -o gpurun_out/r2_scan_v20 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2_ncu_v20.log 2>&1
echo "== bench with the scan forced to variant 20 / 22"; date
for v in 20 22; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant $v | tee gpurun_out/r2_bench_ps_scan_v$v.json
done
echo "== Caduceus-Ph bench under 3 / 12 / 20"; date
for v in 3 12 20; do
  timeout 200 python bench.py --model ph --steps 5 --warmup 3 --no-cpu-baseline --scan-variant $v | tee gpurun_out/r2_bench_ph_scan_v$v.json
done
echo "== full GPU suite with the scan defaulting to variant 11 where it applies"; date
CAD_RUN_UNMEASURED=1 CAD_SCAN_VARIANT=11 timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
date
