#!/bin/bash
# Round-2 GPU call 13: first hardware run of the tcgen05 / TMEM projection kernel (csrc/xproj_umma.cu).
# 1. scripts/umma_probe: the shared-memory descriptor reading (both LBO/SBO assignments), instruction descriptor, TMEM alloc,
#    commit -> mbarrier, tcgen05.ld — one process per case under `timeout`.  2. parity of the kernel (tests/test_gpu_xproj_umma.py).
# 3. timing next to the mma.sync kernel, ncu --set full of the new kernel, bench line with the new kernel.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call13.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== umma probe"
ok0=1; ok1=1
for t in 0 1 2 3; do
  for m in 0 1; do
    timeout 30 ./scripts/_bin/umma_probe $t $m; rc=$?
    echo "probe test $t mode $m rc $rc"
    if [ $m = 0 ] && [ $rc != 0 ]; then ok0=0; fi
    if [ $m = 1 ] && [ $rc != 0 ]; then ok1=0; fi
  done
done
echo "probe summary: mode0_all_ok=$ok0 mode1_all_ok=$ok1"
if [ $ok0 = 0 ] && [ $ok1 = 1 ]; then export CAD_UMMA_DESC_SWAP=1; echo "using CAD_UMMA_DESC_SWAP=1"; fi
if [ $ok0 = 0 ] && [ $ok1 = 0 ]; then echo "NO descriptor reading passed: kernel tests will still run for their diagnostics"; fi
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader || echo "GPU not answering after the probe"
echo "== parity of the kernel"; date
timeout 600 python -m pytest tests/test_gpu_xproj_umma.py -m gpu -q --timeout 200 -x 2>&1 | tail -25
echo "== rest of the file without -x"; date
timeout 600 python -m pytest tests/test_gpu_xproj_umma.py -m gpu -q --timeout 200 2>&1 | tail -15
echo "== timing: mma.sync vs tcgen05"; date
for k in mma umma; do
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 | tee -a gpurun_out/r2c13_xproj_timing.jsonl
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 --model ph | tee -a gpurun_out/r2c13_xproj_timing.jsonl
done
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --bcT 0 | tee -a gpurun_out/r2c13_xproj_timing.jsonl
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --L 16384 | tee -a gpurun_out/r2c13_xproj_timing.jsonl
timeout 120 python scripts/time_xproj.py --kernel mma --iters 20 --L 16384 | tee -a gpurun_out/r2c13_xproj_timing.jsonl
echo "== ncu --set full of the tcgen05 kernel"; date
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_xproj_umma -s 3 -c 1 -f -o gpurun_out/r2c13_xproj_umma \
    python scripts/time_xproj.py --kernel umma --iters 2 > gpurun_out/r2c13_ncu_xproj.log 2>&1
tail -2 gpurun_out/r2c13_ncu_xproj.log
echo "== bench N=1 with the tcgen05 kernel, then default"; date
CAD_XPROJ_KERNEL=umma timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2c13_bench_err.log | grep '^{' | tee gpurun_out/r2c13_bench_ps_umma.json
CAD_XPROJ_KERNEL=mma timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/r2c13_bench_err.log | grep '^{' | tee gpurun_out/r2c13_bench_ps_mma.json
tail -3 gpurun_out/r2c13_bench_err.log | cut -c1-300
date
