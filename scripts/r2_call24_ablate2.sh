#!/bin/bash
# Round-2 GPU call 24: second ablation of the projection kernel's skeleton (diagnostic flags, wrong results by design):
# 1 = no W_x slab loads, 2 = no x slab loads, 8 = no delta stores, 16 = no B / C stores, 32 = no conv arithmetic, 64 = no dt_proj at all
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call24.log) 2>&1
date
timeout 200 python scripts/ab_xproj_flags.py 0 1 2 3 56 57 59 120 123 | tee gpurun_out/r2c24_ablate.json
date
