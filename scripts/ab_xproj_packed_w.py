"""A/B of the conv_xproj kernel with and without the pre-packed W_x operand inside ONE process, interleaved (robust to clock drift
between runs; the same harness timed the experimental switches of profiles/r2_call20..24*)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from caduceus_b200 import functional as CF
dev, L, E, N, R = "cuda", 131072, 512, 16, 16
jobs = CF.job_tables(1, 2, 2, False, torch.device(dev))
g = torch.Generator().manual_seed(0)
sets = [torch.randn(2, 2 * E, L, device=dev, dtype=torch.bfloat16) for _ in range(2)]
w_x = (torch.randn(2, R + 2 * N, E, generator=g) * E ** -0.5).to(dev).bfloat16()
w_dt = (torch.randn(2, E, R, generator=g) * R ** -0.5).to(dev).bfloat16()
conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(dev)
conv_b = (0.1 * torch.randn(2, E, generator=g)).to(dev)
variants = [int(v) for v in (sys.argv[1:] or ["0", "1"])]      # 0 = W_x gathered by tensor-map boxes, 1 = pre-packed W_x (one bulk copy per slab)
w_x_packed = CF.pack_w_x(w_x, R)
tot = {v: 0.0 for v in variants}
rounds = 6
for r in range(rounds + 1):
    for v in variants:
        wxp = w_x_packed if v else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        CF.conv_xproj(sets[0], w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=True, w_x_packed=wxp)
        e0.record()
        for i in range(8):
            CF.conv_xproj(sets[i & 1], w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=True, w_x_packed=wxp)
        e1.record()
        torch.cuda.synchronize()
        if r:
            tot[v] += e0.elapsed_time(e1) / 8
print(json.dumps({"ab_xproj_packed_w_ms": {str(v): round(tot[v] / rounds, 4) for v in variants}}))
