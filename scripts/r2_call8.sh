#!/bin/bash
# Round-2 GPU call 8: full suite (one process per file), bench with 18 PS segments + softplus series, conv_xproj capture.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call8.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== hw_probe AD"; date
timeout 120 ./scripts/_bin/hw_probe 131072 AD
echo "== bench N=1 (default)"; date
timeout 400 python bench.py --steps 10 --warmup 3 | tee gpurun_out/r2c8_bench_ps.json
echo "== bench Ph"; date
timeout 300 python bench.py --model ph --steps 10 --warmup 3 --no-cpu-baseline | tee gpurun_out/r2c8_bench_ph.json
echo "== conv_xproj timing + ncu"; date
timeout 120 python scripts/time_xproj.py | tee gpurun_out/r2c8_xproj.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_xproj -s 3 -c 1 -f -o gpurun_out/r2c8_xproj \
    python scripts/time_xproj.py --iters 2 > gpurun_out/r2c8_ncu_xproj.log 2>&1
tail -1 gpurun_out/r2c8_ncu_xproj.log
echo "== ncu pass A + fix-up at 18 segments"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:bimamba_scan_fwd_v20 -s 24 -c 1 -f \
    -o gpurun_out/r2c8_scan_v20_nseg18 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c8_ncu_v20.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 21 -c 1 -f \
    -o gpurun_out/r2c8_fixup_nseg18 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c8_ncu_fixup.log 2>&1
echo "== full GPU suite, one process per file"; date
bash scripts/gpu_suite_by_file.sh --durations=8
date
