"""Per-layer kernel times of ONE rank of a sequence-sharded Caduceus-PS forward (one GPU, no peers needed): in_proj, conv_xproj,
zero-carry scan with state outputs, whole-shard carry fix-up (random carry-in), out_proj, add+norm — at the shard lengths of the
2- / 4- / 8-way split of 131072 tokens.  CUDA events, 20 iterations after warm-up.
    python scripts/time_shard_pieces.py [--L 16384,32768,65536]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from caduceus_b200 import functional as CF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", default="16384,32768,65536")
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev, E, N, R, D = "cuda", 512, 16, 16, 256
jobs = CF.job_tables(1, 2, 2, False, torch.device(dev))
njobs = jobs[0].numel()
g = torch.Generator().manual_seed(0)
w_x = (torch.randn(2, R + 2 * N, E, generator=g) * E ** -0.5).to(dev).bfloat16()
w_dt = (torch.randn(2, E, R, generator=g) * 0.1).to(dev).bfloat16()
conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(dev)
conv_b = (0.1 * torch.randn(2, E, generator=g)).to(dev)
dt_b = torch.log(torch.expm1(torch.exp(torch.rand(2, E, generator=g) * 4.6 - 6.9))).to(dev)
A2 = (-torch.arange(1, N + 1, dtype=torch.float32).repeat(2, E, 1) * 1.4426950408889634).to(dev).contiguous()
packed = (conv_w4, conv_b, dt_b, A2, torch.ones(2, E, device=dev))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters * 1e3, 1)          # microseconds


for L in [int(v) for v in a.L.split(",")]:
    xz = torch.randn(2, 2 * E, L, device=dev, dtype=torch.bfloat16)
    halo = torch.randn(njobs, E, 3, device=dev, dtype=torch.bfloat16)
    h0 = torch.randn(njobs, E, N, device=dev)
    variant = CF.choose_scan_variant(torch.bfloat16, N, njobs, E, L)
    res = {"L": L, "scan_variant": variant}
    if variant == 20:
        delta, bc, bcT = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, halo=halo, want_bcT=True)
        res["conv_xproj_us"] = timed(lambda: CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, halo=halo, want_bcT=True), a.iters)
    else:
        delta, bc = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, halo=halo)
        bcT = None
        res["conv_xproj_us"] = timed(lambda: CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, halo=halo), a.iters)
    out, hl, ds, ctx = CF.scan_fwd(xz, delta, bc, packed, jobs, L, halo=halo, want_state=True, variant=variant, bcT=bcT)
    res["scan_zero_carry_us"] = timed(lambda: CF.scan_fwd(xz, delta, bc, packed, jobs, L, halo=halo, want_state=True, variant=variant, bcT=bcT), a.iters)
    seg = ctx if isinstance(ctx, dict) else None
    res["fixup_us"] = timed(lambda: CF.scan_fixup(xz, delta, bc, out, packed, jobs, L, h0, seg_ctx=seg), a.iters)
    res["fixup_cut40_us"] = timed(lambda: CF.scan_fixup(xz, delta, bc, out, packed, jobs, L, h0, seg_ctx=seg, cutoff_log2=-40.0), a.iters) if seg is None else None
    hid = torch.randn(L, 2 * D, device=dev, dtype=torch.bfloat16)
    w_in = torch.randn(2 * E, D, device=dev, dtype=torch.bfloat16)
    w_out = torch.randn(2 * E, D, device=dev, dtype=torch.bfloat16)
    res["in_proj_2strands_us"] = timed(lambda: [torch.mm(w_in, hid[:, s * D:(s + 1) * D].t()) for s in range(2)], a.iters)
    yg = torch.randn(2, 2 * E, L, device=dev, dtype=torch.bfloat16)
    res["out_proj_2strands_us"] = timed(lambda: [torch.mm(yg[s].t(), w_out) for s in range(2)], a.iters)
    wn = torch.ones(D, device=dev, dtype=torch.bfloat16)
    res["add_norm_us"] = timed(lambda: CF.add_norm(hid[None], wn, None, residual=hid[None], eps=1e-5, is_rms=True, prenorm=True,
                                                   nhalf=2, swap=1, wflip_mask=2), a.iters)
    print(json.dumps(res), flush=True)
