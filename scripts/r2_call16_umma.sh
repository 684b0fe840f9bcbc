#!/bin/bash
# Round-2 GPU call 16: tcgen05 projection kernel with hardware-parked waits and x slabs requested a full ring ahead; the kernel
# as the default of the model path: parity files that exercise it, bench line.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call16.log) 2>&1
date
echo "== parity of the kernel"
timeout 600 python -m pytest tests/test_gpu_xproj_umma.py -m gpu -q --timeout 200 2>&1 | tail -15
echo "== timing"; date
for k in umma mma; do
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 | tee -a gpurun_out/r2c16_xproj_timing.jsonl
done
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --model ph | tee -a gpurun_out/r2c16_xproj_timing.jsonl
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --L 16384 | tee -a gpurun_out/r2c16_xproj_timing.jsonl
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --bcT 0 | tee -a gpurun_out/r2c16_xproj_timing.jsonl
echo "== ncu --set full"; date
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_xproj_umma -s 3 -c 1 -f -o gpurun_out/r2c16_xproj_umma \
    python scripts/time_xproj.py --kernel umma --iters 2 > gpurun_out/r2c16_ncu_xproj.log 2>&1
tail -1 gpurun_out/r2c16_ncu_xproj.log
echo "== model-level parity with the kernel as default"; date
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scan_variants.py -m gpu -q --timeout 300 -k "xproj or headline or fixture or model or mixer or graphed" 2>&1 | tail -8
echo "== bench N=1"; date
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2c16_bench_err.log | grep '^{' | tee gpurun_out/r2c16_bench_ps.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --model ph 2>>gpurun_out/r2c16_bench_err.log | grep '^{' | tee gpurun_out/r2c16_bench_ph.json
tail -3 gpurun_out/r2c16_bench_err.log | cut -c1-300
date
