"""Time the same-box Blackwell kernel to beat: vLLM's sm_100 build of the UPSTREAM selective_scan_fwd CUDA kernel
(vllm._C.selective_scan_fwd, derived from mamba_ssm's csrc) on the work ONE of our bidirectional launches covers.

The reference (ref:caduceus/modeling_caduceus.py:122-140, ref:caduceus/modeling_rcps.py:80-101) issues, per BiMamba call and
strand, one scan on the sequence and one on the flipped sequence, i.e. per Caduceus-PS layer: 4 selective-scan launches at
B = 1, E = 512, N = 16 plus 2 input flips and 2 output flips (the RC strand adds its own flips; not counted here, and upstream's
separate causal_conv1d launch is not counted either: both omissions favour the upstream number).

    python scripts/time_vllm_scan.py [--L 131072] [--iters 10]
One JSON line per model (ps / ph): upstream ms per PS-equivalent with and without flips; ours is printed by
scripts/time_scan_variants.py on the same box."""
import argparse
import json

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=131072)
ap.add_argument("--E", type=int, default=512)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--dtype", default="bf16")
args = ap.parse_args()

from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn  # noqa: E402

dev = "cuda"
L, E, N = args.L, args.E, 16
dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
g = torch.Generator(device="cpu").manual_seed(0)


def make_set():
    u = torch.randn(1, E, L, device=dev).to(dt)
    delta = torch.randn(1, E, L, device=dev).to(dt)
    z = torch.randn(1, E, L, device=dev).to(dt)
    B = torch.randn(1, 1, N, L, device=dev).to(dt)
    C = torch.randn(1, 1, N, L, device=dev).to(dt)
    return u, delta, z, B, C


A = -torch.arange(1, N + 1, dtype=torch.float32, device=dev).repeat(E, 1).contiguous()
D = torch.ones(E, device=dev)
dt_b = torch.log(torch.expm1(torch.exp(torch.rand(E, generator=g) * 4.6 - 6.9))).to(dev)

for model, nstrand in (("ps", 2), ("ph", 1)):
    # rotating sets so that no launch finds its inputs in L2 (each set: 3 x 128 MiB + B/C)
    sets = [make_set() for _ in range(3)]

    def one_direction(s, flip):
        u, delta, z, B, C = s
        if flip:   # the reference flips the hidden states before in_proj; here the cheapest equivalent: flip the scan inputs' source once
            u = u.flip(-1)
        states = torch.zeros(1, E, N, device=dev)
        out = selective_scan_fn(u, states, delta, A, B, C, D, z, dt_b, delta_softplus=True)   # writes its output over z
        if flip:
            out = out.flip(-1)
        return out

    def step(i, flips):
        for sidx in range(nstrand):
            s = sets[(i + sidx) % len(sets)]
            one_direction(s, False)
            one_direction(s, flips)

    res = {}
    for flips in (False, True):
        for _ in range(3):
            step(0, flips)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            step(i, flips)
        e1.record()
        torch.cuda.synchronize()
        res["with_flips" if flips else "scan_only"] = round(e0.elapsed_time(e1) / args.iters, 4)
    print(json.dumps({"kernel": "vllm._C.selective_scan_fwd (upstream mamba_ssm kernel, sm_100 build)", "model": model,
                      "L": L, "E": E, "dtype": args.dtype, "launches_per_step": 2 * nstrand,
                      "ms_per_bimamba_equivalent": res}), flush=True)
