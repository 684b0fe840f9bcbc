#!/bin/bash
# Round-2 GPU call 15: warp-specialised version of the tcgen05 projection kernel (TMA x slabs, resident taps / W_dt, channels on the TMEM
# lanes for dt_proj): parity, timing next to the mma.sync kernel, ncu --set full.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call15.log) 2>&1
date
echo "== parity of the kernel"
timeout 600 python -m pytest tests/test_gpu_xproj_umma.py -m gpu -q --timeout 200 2>&1 | tail -25
echo "== timing: mma.sync vs tcgen05"; date
for k in mma umma; do
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 | tee -a gpurun_out/r2c15_xproj_timing.jsonl
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 --model ph | tee -a gpurun_out/r2c15_xproj_timing.jsonl
  timeout 120 python scripts/time_xproj.py --kernel $k --iters 20 --L 16384 | tee -a gpurun_out/r2c15_xproj_timing.jsonl
done
timeout 120 python scripts/time_xproj.py --kernel umma --iters 20 --bcT 0 | tee -a gpurun_out/r2c15_xproj_timing.jsonl
echo "== ncu --set full of the tcgen05 kernel"; date
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_xproj_umma -s 3 -c 1 -f -o gpurun_out/r2c15_xproj_umma \
    python scripts/time_xproj.py --kernel umma --iters 2 > gpurun_out/r2c15_ncu_xproj.log 2>&1
tail -2 gpurun_out/r2c15_ncu_xproj.log
date
