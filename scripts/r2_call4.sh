#!/bin/bash
# Round-2 GPU call 4: pass A with word-wise x / z reads, fix-up at 6 CTAs of 4 warps per SM.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call4.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== hw_probe D"; date
timeout 120 ./scripts/_bin/hw_probe 131072 D
echo "== fix-up + variant tests on hardware"; date
CAD_RUN_UNMEASURED=1 timeout 600 python -m pytest tests/test_gpu_scan_variants.py tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "fixup or v20 or segment or shard or variant" 2>&1 | tail -4
echo "== ncu v20 pass A, nseg 37 W 8"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:bimamba_scan_fwd_v20 -s 7 -c 1 -f \
    -o gpurun_out/r2c4_scan_v20_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c4_ncu_v20.log 2>&1
tail -1 gpurun_out/r2c4_ncu_v20.log
echo "== ncu fix-up, nseg 37"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 0 -c 1 -f \
    -o gpurun_out/r2c4_fixup_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c4_ncu_fixup.log 2>&1
tail -1 gpurun_out/r2c4_ncu_fixup.log
date
