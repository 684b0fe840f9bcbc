#!/bin/bash
# Round-2 GPU call 21: ablation timing of the tcgen05 projection kernel (diagnostic flags, wrong results by design): what bounds it?
# 8 = no delta stores, 16 = no B / C stores, 32 = no conv arithmetic (raw x copied into the operand)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call21.log) 2>&1
date
timeout 200 python scripts/ab_xproj_flags.py 0 8 16 24 32 40 56 | tee gpurun_out/r2c21_ablate.json
date
