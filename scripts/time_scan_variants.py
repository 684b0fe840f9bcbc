"""A/B timing of the two forward-scan kernels (3: time-parallel, 20: lane = channel pipeline) on the headline shapes (one GPU):
    python scripts/time_scan_variants.py [--model ps|ph] [--L 131072] [--iters 10]
Prints one JSON line per variant: ms per launch (CUDA events on the launching stream, after warm-up), boundary-S
GB/s (8320 B per nucleotide per BiMamba call) and the max abs difference to variant 3's output."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from caduceus_b200 import functional as CF  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="ps", help="ps, ph or ps,ph")
ap.add_argument("--L", type=int, default=131072)
ap.add_argument("--E", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--variants", default="3,20")
args = ap.parse_args()

dev = "cuda"
for model in args.model.split(","):
    L, E, N = args.L, args.E, 16
    nstrand = 2 if model == "ps" else 1
    njobs = 2 * nstrand
    g = torch.Generator(device="cpu").manual_seed(0)
    ld = CF.round_up(L, 16)
    ldbc = CF.round_up(L, 32)
    # two rotating input sets (each > 1 GB for PS) so that no launch finds its inputs in the 126 MB L2
    sets = []
    for r in range(2):
        xz = torch.randn(nstrand, 2 * E, ld, device=dev, dtype=torch.bfloat16)
        delta = (torch.randn(njobs, E, ld, device=dev) * 1.0).to(torch.bfloat16)
        bc = torch.zeros(njobs, 2 * N, ldbc, device=dev)
        bc[..., :L] = torch.randn(njobs, 2 * N, L, device=dev)
        bc[..., :L] = bc[..., :L].bfloat16().float()          # bf16-representable B / C values, as the model's x_dbl
        sets.append((xz, delta, bc))
    conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(dev)
    conv_b = (0.1 * torch.randn(2, E, generator=g)).to(dev)
    dt_b = torch.log(torch.expm1(torch.exp(torch.rand(2, E, generator=g) * 4.6 - 6.9))).to(dev)
    A2 = (-torch.arange(1, N + 1, dtype=torch.float32).repeat(2, E, 1) * 1.4426950408889634).to(dev).contiguous()
    Dk = torch.ones(2, E, device=dev)
    packed = (conv_w4, conv_b, dt_b, A2, Dk)
    # PS job order: (strand, direction); rev = direction XOR strand.  Ph: (fwd, rev)
    seq = torch.tensor([j // 2 for j in range(njobs)], dtype=torch.int32, device=dev)
    pset = torch.tensor([j % 2 for j in range(njobs)], dtype=torch.int32, device=dev)
    rev = torch.tensor([(j % 2) ^ (j // 2) for j in range(njobs)], dtype=torch.int32, device=dev)
    jobs = (seq, pset, rev)

    base = None
    for v in [int(s) for s in args.variants.split(",")]:
        try:
            out = None
            for _ in range(3):
                out, _, _, _ = CF.scan_fwd(*sets[0][:3], packed, jobs, L, variant=v)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.iters):
                CF.scan_fwd(*sets[i & 1][:3], packed, jobs, L, variant=v)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            o = out[..., :L].float()
            if base is None:
                base = o
            nt_calls = nstrand * L
            print(json.dumps({"variant": v, "model": model, "L": L, "E": E, "ms_per_launch": round(ms, 4),
                              "boundary_S_GBps": round(8320 * nt_calls / ms / 1e6, 1),
                              "finite": bool(torch.isfinite(o).all()), "max_abs": float(o.abs().max()),
                              "max_abs_diff_vs_first": float((o - base).abs().max())}), flush=True)
        except Exception as ex:  # keep going: the other variant's number is still wanted
            print(json.dumps({"variant": v, "error": repr(ex)[:400]}), flush=True)
