#!/bin/bash
# Round-2 GPU call 2: pipe rates (FFMA2 / MUFU f16x2 / mixes) + ncu --set full of v20 pass A at 37 segments and of the segment fix-up.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call2.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pipe_probe"; date
timeout 120 ./scripts/_bin/pipe_probe | tee gpurun_out/r2_pipe_probe.log
echo "== ncu v20 pass A, nseg 37 W 8"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:bimamba_scan_fwd_v20 -s 7 -c 1 -f \
    -o gpurun_out/r2_scan_v20_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2_ncu_v20b.log 2>&1
tail -2 gpurun_out/r2_ncu_v20b.log
echo "== ncu fix-up, nseg 37"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 0 -c 1 -f \
    -o gpurun_out/r2_fixup_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2_ncu_fixup.log 2>&1
tail -2 gpurun_out/r2_ncu_fixup.log
date
