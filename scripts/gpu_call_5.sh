#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/call5.log) 2>&1
date
timeout 120 python scripts/time_scan_variants.py --model ps,ph --variants 3,7,8 | tee gpurun_out/ab_scan5.jsonl
for v in 7 8; do
  echo "== scan-level + model parity with CAD_SCAN_VARIANT=$v"; date
  CAD_SCAN_VARIANT=$v timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q --timeout 60 \
      -k "scan or mixer or block or model or full_length or fixup" 2>&1 | tail -4
done
date
