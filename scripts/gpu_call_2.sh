#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/call2.log) 2>&1
date
echo "== A/B timing (ps, ph): variants 3 5 6 4"
timeout 200 python scripts/time_scan_variants.py --model ps,ph --variants 3,5,6,4 | tee gpurun_out/ab_scan2.jsonl
for v in 5 6; do
  echo "== scan-level + model parity with CAD_SCAN_VARIANT=$v"; date
  CAD_SCAN_VARIANT=$v timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q --timeout 120 \
      -k "scan or mixer or block or model or full_length or fixup" 2>&1 | tail -6
done
date
