#!/bin/bash
# Round-2 GPU call 6: pruned library (variants 3 + 20), fix-up with L1-allocating cp.async ring, full GPU suite, bench.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call6.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== hw_probe AD"; date
timeout 120 ./scripts/_bin/hw_probe 131072 AD
echo "== full GPU suite"; date
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25
echo "== bench N=1 (default)"; date
timeout 400 python bench.py --steps 5 --warmup 3 | tee gpurun_out/r2c6_bench_ps.json
echo "== bench N=1 forced v3"; date
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --scan-variant 3 | tee gpurun_out/r2c6_bench_ps_v3.json
echo "== bench Ph"; date
timeout 300 python bench.py --model ph --steps 5 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2c6_bench_ph.json
echo "== ncu fix-up, nseg 37"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 0 -c 1 -f \
    -o gpurun_out/r2c6_fixup_nseg37 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c6_ncu_fixup.log 2>&1
tail -1 gpurun_out/r2c6_ncu_fixup.log
date
