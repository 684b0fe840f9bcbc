#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/call3.log) 2>&1
date
ST="0,100,200,300,400,600,65736,65836,65936,131272,131372"   # mode 0: 100..600; mode 1: 200,300,400; mode 2: 200,300
timeout 300 python scripts/time_scan_variants.py --model ps,ph --variants 3,5 --staggers $ST | tee gpurun_out/ab_scan3.jsonl
date
