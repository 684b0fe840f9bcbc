#!/bin/bash
# Round-2 GPU call 19: final single-GPU validation of the round's end state — smoke, the -m gpu suite in ONE process exactly as the
# driver runs it, bench (both arms), launch list of a 2-layer PS forward, ncu --set full of the tcgen05 projection kernel.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call19.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; date
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
echo "== bench N=1 (default)"; date
timeout 400 python bench.py | tee gpurun_out/r2c19_bench_ps.json
echo "== bench --impl reference"; date
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/r2c19_bench_reference.json
echo "== launch list of a 2-layer PS forward"; date
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c19_launches_ps_2layer.csv \
    python bench.py --steps 1 --warmup 1 --n-layer 2 --no-cpu-baseline > gpurun_out/r2c19_launches.log 2>&1
tail -2 gpurun_out/r2c19_launches.log | cut -c1-200
echo "== ncu --set full of the tcgen05 projection kernel (final source)"; date
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_xproj_umma -s 3 -c 1 -f -o gpurun_out/r2c19_xproj_umma \
    python scripts/time_xproj.py --iters 2 > gpurun_out/r2c19_ncu_xproj.log 2>&1
tail -1 gpurun_out/r2c19_ncu_xproj.log
echo "== the -m gpu suite, one process, as the driver runs it"; date
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 --durations=6 2>&1 | tail -25
date
