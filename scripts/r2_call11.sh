#!/bin/bash
# Round-2 GPU call 11: final single-GPU validation — smoke, full suite (one process per file), bench (both arms), fix-up capture at 18 segments.
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call11.log) 2>&1
date; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke"; date
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
echo "== full GPU suite, one process per file"; date
bash scripts/gpu_suite_by_file.sh --durations=4
echo "== bench N=1 (default)"; date
timeout 400 python bench.py | tee gpurun_out/r2c11_bench_ps.json
echo "== bench --impl reference"; date
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/r2c11_bench_reference.json
echo "== ncu fix-up at 18 segments (cut -16)"; date
timeout 300 ncu --set full --clock-control none --import-source on --target-processes all -k regex:scan_fixup -s 39 -c 1 -f \
    -o gpurun_out/r2c11_fixup_nseg18 ./scripts/_bin/hw_probe 131072 D > gpurun_out/r2c11_ncu_fixup.log 2>&1
tail -1 gpurun_out/r2c11_ncu_fixup.log
echo "== launch list of a 2-layer PS forward"; date
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c11_launches_ps_2layer.csv \
    python bench.py --steps 1 --warmup 1 --n-layer 2 --no-cpu-baseline > gpurun_out/r2c11_launches.log 2>&1
tail -2 gpurun_out/r2c11_launches.log | cut -c1-200
date
