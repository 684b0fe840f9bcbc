#!/bin/bash
# BASELINE.json configs[4]: seq_len sweep 1k/8k/65k/131k/524k at d_model=256, n_layer=16, bf16 forward, 1 x B200.
# usage (on the GPU box): bash profiles/seqlen_sweep.sh > gpurun_out/seqlen_sweep.jsonl
for model in ps ph; do
  for L in 1024 8192 65536 131072 524288; do
    python bench.py --model $model --seqlen $L --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
  done
done
