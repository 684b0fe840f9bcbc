#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call6b.log) 2>&1
date
bash scripts/gpu_suite_by_file.sh -x
date
