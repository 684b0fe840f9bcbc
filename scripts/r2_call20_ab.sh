#!/bin/bash
# Round-2 GPU call 20: interleaved A/B of back-off sleeps in the waits of the idle roles of the tcgen05 projection kernel.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call20.log) 2>&1
date
timeout 200 python scripts/ab_xproj_flags.py 0 1 2 3 5 7 | tee gpurun_out/r2c20_ab.json
timeout 200 python scripts/ab_xproj_flags.py 0 3 7 | tee -a gpurun_out/r2c20_ab.json
timeout 300 python -m pytest tests/test_gpu_xproj.py -m gpu -q --timeout 200 2>&1 | tail -2
date
