"""Time (and give ncu something to capture of) the fused conv + SiLU -> x_proj -> dt_proj kernel on the headline shapes:
    python scripts/time_xproj.py [--model ps|ph] [--L 131072] [--iters 10] [--bcT 1]
One JSON line: ms per launch (CUDA events) and the HBM floor of its algorithmic traffic (read x, write delta + bc [+ bcT])."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from caduceus_b200 import functional as CF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ps")
ap.add_argument("--L", type=int, default=131072)
ap.add_argument("--E", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--bcT", type=int, default=1)
a = ap.parse_args()
dev, L, E, N, R = "cuda", a.L, a.E, 16, 16
nstrand = 2 if a.model == "ps" else 1
jobs = CF.job_tables(1, nstrand, 2, False, torch.device(dev))
njobs = jobs[0].numel()
g = torch.Generator().manual_seed(0)
sets = [torch.randn(nstrand, 2 * E, CF.round_up(L, 16), device=dev, dtype=torch.bfloat16) for _ in range(2)]
w_x = (torch.randn(2, R + 2 * N, E, generator=g) * E ** -0.5).to(dev).bfloat16()
w_dt = (torch.randn(2, E, R, generator=g) * R ** -0.5).to(dev).bfloat16()
conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(dev)
conv_b = (0.1 * torch.randn(2, E, generator=g)).to(dev)
for _ in range(3):
    CF.conv_xproj(sets[0], w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=bool(a.bcT))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.iters):
    CF.conv_xproj(sets[i & 1], w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=bool(a.bcT))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
bytes_ = njobs * E * L * 2 * 2 + njobs * 2 * N * L * 4 * (2 if a.bcT else 1)
print(json.dumps({"kernel": "conv_xproj_umma_kernel", "model": a.model, "L": L, "E": E, "bcT": bool(a.bcT), "ms_per_launch": round(ms, 4),
                  "algorithmic_GB": round(bytes_ / 1e9, 3), "GBps": round(bytes_ / ms / 1e6, 1)}))
