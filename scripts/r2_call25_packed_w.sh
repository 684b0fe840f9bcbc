#!/bin/bash
# Round-2 GPU call 25 (last seconds of the budget): W_x handed to the projection kernel pre-packed per slab (one bulk copy instead of
# eight tensor-map boxes of 16-byte rows): parity of both paths, then interleaved A/B timing.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call25.log) 2>&1
date
timeout 70 python -m pytest tests/test_gpu_xproj.py -m gpu -q -x --timeout 60 2>&1 | tail -4
timeout 40 python scripts/ab_xproj_packed_w.py 0 1 | tee gpurun_out/r2c25_ab.json
date
