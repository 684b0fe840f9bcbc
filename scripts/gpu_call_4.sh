#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/call4.log) 2>&1
date
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bimamba_scan_fwd_kernel -s 2 -c 1 -f -o gpurun_out/scan_v3 \
    python scripts/time_scan_variants.py --model ps --variants 3 --iters 1 2>&1 | tail -5
ls -la gpurun_out | tail -3
date
