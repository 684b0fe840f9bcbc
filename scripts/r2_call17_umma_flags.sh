#!/bin/bash
# Round-2 GPU call 17: which of the two call-16 changes cost time?  CAD_UMMA_FLAGS bit 0: role warps wait with a suspend hint,
# bit 1: conv warps too, bit 2: x slab requests also wait for the MMAs of the slot (the call-15 gating).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call17.log) 2>&1
date
for f in 0 4 1 2 3 5; do
  echo "flags $f"
  CAD_UMMA_FLAGS=$f timeout 120 python scripts/time_xproj.py --kernel umma --iters 40 | tee -a gpurun_out/r2c17_xproj_flags.jsonl
done
CAD_UMMA_FLAGS=0 timeout 120 python scripts/time_xproj.py --kernel umma --iters 40 --model ph
CAD_UMMA_FLAGS=4 timeout 120 python scripts/time_xproj.py --kernel umma --iters 40 --model ph
date
