#!/bin/bash
# Round-2 GPU call 5: fix-up with a cp.async ring for the C rows (8 CTAs of 4 warps per SM), pass A back to per-token smem reads;
# first hardware run of the peer-exchange kernels (virtual ranks on one GPU).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call5.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== hw_probe D"; date
timeout 120 ./scripts/_bin/hw_probe 131072 D
echo "== peer exchange, virtual ranks"; date
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 -x 2>&1 | tail -12
echo "== fix-up + variant tests on hardware"; date
CAD_RUN_UNMEASURED=1 timeout 600 python -m pytest tests/test_gpu_scan_variants.py tests/test_gpu_parity.py -m gpu -q --timeout 120 -x -k "fixup or v20 or segment or shard or variant" 2>&1 | tail -4
echo "== ncu fix-up, nseg 37"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 0 -c 1 -f \
    -o gpurun_out/r2c5_fixup_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c5_ncu_fixup.log 2>&1
tail -1 gpurun_out/r2c5_ncu_fixup.log
date
