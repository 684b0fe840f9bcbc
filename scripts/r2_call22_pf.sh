#!/bin/bash
# Round-2 GPU call 22: L2 tensor prefetch of the x slabs `pf` slabs ahead of the shared-memory ring (CAD_UMMA_FLAGS = pf << 8).
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call22.log) 2>&1
date
timeout 200 python scripts/ab_xproj_flags.py 0 1024 1536 2048 3072 4096 6144 | tee gpurun_out/r2c22_pf.json
CAD_UMMA_FLAGS=2048 timeout 300 python -m pytest tests/test_gpu_xproj.py -m gpu -q --timeout 200 2>&1 | tail -2
date
