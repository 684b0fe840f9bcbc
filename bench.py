#!/usr/bin/env python
"""Headline benchmark: nucleotides/s of the Caduceus forward at seq_len=131072, d_model=256, n_layer=16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model ps|ph] [--seqlen L] [--impl reference]

Workload (BASELINE.json configs[2]): Caduceus-PS (rcps=true) d_model=256 n_layer=16 seq_len=131072 bf16
forward, batch 1 per GPU, random-init weights (reference init), synthetic hg38-shaped ids.
A "step" = one forward over one batch.  N > 1: one process per GPU (torchrun), every rank runs its own sequence
(the reference's own scaling mode is data-parallel, ref:train.py:629-639) -> "scaling": "weak"; there is no
data-path collective, only the timing barrier.

`--impl reference` times the reference's CPU path for the same metric: the oracle port of upstream
`selective_scan_ref` (oracle/mamba_ssm/ops/selective_scan_interface.py) — the CPU implementation BASELINE.json
names — on a bounded sample, on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "nucleotides/sec at seq_len=131072, d_model=256, n_layer=16 (whole job)"
UNIT = "nt/s"
CMAP = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 10, 8: 9, 9: 8, 10: 7, 11: 11}
SSM_CFG = dict(d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1, dt_init="random",
               dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="ps", choices=["ps", "ph"])
    ap.add_argument("--seqlen", type=int, default=131072)
    ap.add_argument("--d-model", type=int, default=256)
    ap.add_argument("--n-layer", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="forward", choices=["forward", "train"],
                    help="train: MLM step = forward + loss + backward + AdamW under bf16 autocast (BASELINE configs[3] on 1 GPU)")
    ap.add_argument("--graph", action="store_true",
                    help="replay the forward from a CUDA graph (single-GPU / replica forward only; helps short sequences)")
    ap.add_argument("--scan-tok", type=int, default=0, choices=[0, 8, 16], help="tokens per lane of the scan kernel (tuning)")
    ap.add_argument("--scan-variant", type=int, default=None, choices=[0, 3, 4, 7, 9, 10, 11, 12, 20, 21, 22, 23],
                    help="forward-scan kernel variant (cad_scan_fwd_args.variant); default: env CAD_SCAN_VARIANT / library default")
    ap.add_argument("--shard", default="none", choices=["none", "seq"],
                    help="seq: ONE sequence of --seqlen sharded on the sequence axis over all ranks (strong scaling)")
    return ap.parse_args()


def workload_name(a):
    what = "forward" if a.mode == "forward" else "MLM train step (fwd+bwd+AdamW, autocast)"
    return (f"Caduceus-{'PS (rcps=true)' if a.model == 'ps' else 'Ph'} d_model={a.d_model} n_layer={a.n_layer} "
            f"seq_len={a.seqlen} bf16 {what}, batch {a.batch}/GPU")


def scans_per_nt(a):
    """token-direction scans one nucleotide goes through in a forward: layers x directions x strands."""
    return a.n_layer * 2 * (2 if a.model == "ps" else 1)


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of selective_scan_ref
# ---------------------------------------------------------------------------------------------------------
def cpu_scan_rate(seconds_budget, L=2048, E=512, N=16, min_calls=1):
    """token-directions/s of the restated selective_scan_ref (fp32, all host threads), bounded sample."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mamba_ssm.ops.selective_scan_interface import selective_scan_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    u, z = torch.randn(1, E, L, generator=g), torch.randn(1, E, L, generator=g)
    Bm, Cm = torch.randn(1, N, L, generator=g), torch.randn(1, N, L, generator=g)
    delta = 0.5 * torch.rand(1, E, L, generator=g)
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(E, 1)
    D, dbias = torch.ones(E), 0.5 * torch.rand(E, generator=g)
    calls, t_total = 0, 0.0
    while calls < min_calls or t_total < seconds_budget:
        t0 = time.perf_counter()
        selective_scan_ref(u, delta, A, Bm, Cm, D, z, dbias, delta_softplus=True)
        t_total += time.perf_counter() - t0
        calls += 1
    return calls * L / t_total, cores, f"{calls} x selective_scan_ref(B=1, E={E}, N={N}, L={L}) fp32, {cores} threads"


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step, rates = [], []
    sample = ""
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        rate, cores, sample = cpu_scan_rate(0.0, L=2048)
        dt = time.perf_counter() - t0
        if i >= a.warmup:
            per_step.append(dt)
            rates.append(rate)
    td_per_s = sum(rates) / len(rates)
    value = td_per_s / scans_per_nt(a)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * sum(per_step) / len(per_step), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(a),
                   "note": "CPU oracle port of upstream selective_scan_ref (scan only, no projections/conv/norm); "
                           f"nt/s = token-directions/s / {scans_per_nt(a)} scans per nucleotide"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"per step: {sample}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def make_ids(torch, batch, L, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(7, 11, (batch, L), generator=g)
    ids[torch.rand(batch, L, generator=g) < 0.005] = 4        # N -> [PAD], ~0.5 %
    return ids


def run_b200(a):
    import torch
    import torch.distributed as dist
    import caduceus
    from caduceus_b200 import functional as CF
    if a.scan_variant is not None:
        CF.SCAN_VARIANT = a.scan_variant

    CF.SCAN_TOKENS_PER_LANE = a.scan_tok
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    cfg = caduceus.CaduceusConfig(
        d_model=a.d_model, n_layer=a.n_layer, vocab_size=12, ssm_cfg=dict(SSM_CFG), rms_norm=True,
        fused_add_norm=True, residual_in_fp32=False, pad_vocab_size_multiple=8, norm_epsilon=1e-5,
        initializer_cfg=dict(initializer_range=0.02, rescale_prenorm_residual=True, n_residuals_per_layer=1),
        bidirectional=True, bidirectional_strategy="add", bidirectional_weight_tie=True, rcps=(a.model == "ps"),
        complement_map=dict(CMAP) if a.model == "ps" else None)
    train = a.mode == "train"
    if train:
        cfg.pad_token_id = 4
        model = caduceus.CaduceusForMaskedLM(cfg).to(dev).train()          # fp32 master weights, bf16 autocast
        opt = torch.optim.AdamW(model.parameters(), lr=8e-3, weight_decay=0.1, fused=True)
    else:
        model = caduceus.CaduceusForMaskedLM(cfg).to(dev).to(torch.bfloat16).eval()

    # a ring of distinct input batches in pinned host memory; each step's activations (> 1 GB of xz / scan
    # buffers at L=131072) exceed the 126 MB L2, so no explicit flush is needed between iterations
    nbuf = 4
    host_ids = [make_ids(torch, a.batch, a.seqlen, 100 + rank * nbuf + i).pin_memory() for i in range(nbuf)]
    dev_ids = [h.to(dev) for h in host_ids]
    host_out = torch.empty(a.batch, a.seqlen, cfg.vocab_size, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    shard_seq = a.shard == "seq" and world > 1
    if shard_seq:       # every rank holds tokens [rank*L/P, (rank+1)*L/P) of the SAME sequences
        from caduceus_b200 import seqshard
        Ls = a.seqlen // world
        host_ids = [make_ids(torch, a.batch, a.seqlen, 100 + i)[:, rank * Ls:(rank + 1) * Ls].contiguous().pin_memory()
                    for i in range(nbuf)]
        dev_ids = [h.to(dev) for h in host_ids]
        host_out = torch.empty(a.batch, Ls, cfg.vocab_size, dtype=torch.float32).pin_memory()
        _model = model

        def model(ids):          # noqa: F811  (sequence-parallel wrapper around the same module)
            with seqshard.sequence_parallel():
                return _model(ids)

    def mlm(ids):        # 15 % positions: 80 % [MASK], 10 % random, 10 % kept; target PAD elsewhere
        g = torch.Generator(device=ids.device).manual_seed(1)
        r = torch.rand(ids.shape, device=ids.device, generator=g)
        sel = r < 0.15
        inp = ids.clone()
        inp[sel & (r < 0.12)] = 3
        rnd = sel & (r >= 0.12) & (r < 0.135)
        inp[rnd] = torch.randint(0, 12, (int(rnd.sum()),), device=ids.device, generator=g)
        tgt = torch.where(sel, ids, torch.full_like(ids, 4))
        return inp, tgt

    host_loss = torch.empty((), dtype=torch.float32).pin_memory()

    def train_step(ids):
        inp, tgt = mlm(ids)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if shard_seq:
                # every rank: its shard's logits -> CE summed over its masked tokens / the GLOBAL count; the backward
                # gathers the adjoint carries (seqshard.py), then the partial parameter gradients are summed
                logits = model(inp).logits
                cnt = (tgt != 4).sum()
                dist.all_reduce(cnt)
                loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), tgt.reshape(-1),
                                                         ignore_index=4, reduction="sum") / cnt
            else:
                loss = model(inp, labels=tgt).loss
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if shard_seq:
            seqshard.all_reduce_grads(_model)
        opt.step()
        return loss

    def step_device(i):
        if train:
            return train_step(dev_ids[i % nbuf])
        with torch.no_grad():
            return model(dev_ids[i % nbuf]).logits

    def step_e2e(i):
        ids = host_ids[i % nbuf].to(dev, non_blocking=True)
        if train:
            loss = train_step(ids)
            host_loss.copy_(loss.detach(), non_blocking=True)
            return loss
        with torch.no_grad():
            logits = model(ids).logits
            host_out.copy_(logits, non_blocking=True)
        return logits

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        barrier()
        return ms

    for i in range(a.warmup):
        step_device(i)
        step_e2e(i)

    graphed = False
    if a.graph and not train and not shard_seq:
        from caduceus_b200.graphs import GraphedForward
        CF.LAUNCHES = 0
        step_device(0)
        launches_per_step = CF.LAUNCHES
        gfwd = GraphedForward(model, dev_ids[0])
        graphed = True

        def step_device(i):        # noqa: F811
            return gfwd(dev_ids[i % nbuf])

        def step_e2e(i):           # noqa: F811
            logits = gfwd(host_ids[i % nbuf])
            host_out.copy_(logits, non_blocking=True)
            return logits

        for i in range(2):
            step_device(i)
            step_e2e(i)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    CF.LAUNCHES = 0
    CF.SCAN_EVENTS = None if graphed else []     # (start, end) CUDA events around every fused-scan launch
    ms_dev = timed(step_device, a.steps)
    launches = launches_per_step * a.steps if graphed else CF.LAUNCHES     # a replayed graph re-runs the captured launches
    scan_events, CF.SCAN_EVENTS = (CF.SCAN_EVENTS or []), None
    ms_e2e = timed(step_e2e, a.steps)
    if sampler:
        sampler.stop_flag.set()
        sampler.join()

    nt_per_step = a.batch * a.seqlen * (1 if shard_seq else world)
    value = nt_per_step * a.steps / (ms_dev * 1e-3)
    e2e_value = nt_per_step * a.steps / (ms_e2e * 1e-3)

    if rank == 0:
        # roofline of the dominant kernel (fused bidirectional scan), boundary S of SURVEY.md §8d:
        # 2*(8E+4N) B per nucleotide per BiMamba call; a PS launch covers both strands = 2 calls.
        E, N = 2 * a.d_model, 16
        calls_per_launch = 2 if a.model == "ps" else 1
        bytes_per_launch = 2 * (8 * E + 4 * N) * calls_per_launch * a.batch * a.seqlen
        scan_ms = [s.elapsed_time(e) for s, e in scan_events]
        scan_v4 = CF.SCAN_VARIANT == 4 and not train and not shard_seq      # v4 covers the plain inference call only
        avg_scan_ms = sum(scan_ms) / max(len(scan_ms), 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        achieved = bytes_per_launch / (avg_scan_ms * 1e-3) / 1e9 if scan_ms else None
        traffic = None        # measured DRAM bytes per launch, from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
            if a.seqlen == 131072 and a.d_model == 256 and a.batch == 1 and not shard_seq:
                # measured traffic exists per kernel variant: no entry for the running variant -> null, not another kernel's number
                v_run = 4 if scan_v4 else (CF.SCAN_VARIANT if (CF.SCAN_VARIANT and not train) else 3)
                traffic = tj.get(a.model + ("" if v_run == 3 else f"_v{v_run}"), {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        plain = not train and not shard_seq           # the non-default variants cover the plain inference call
        kname = ("bimamba_scan_fwd_v4_kernel" if scan_v4 else
                 "bimamba_scan_fwd_v9_kernel" if CF.SCAN_VARIANT in (9, 10) and plain else
                 "bimamba_scan_fwd_v11_kernel" if CF.SCAN_VARIANT in (11, 12) and plain else
                 "bimamba_scan_fwd_v20_kernel (+ bc_transpose, seg_carry, scan_fixup: the whole segmented scan is timed)"
                 if CF.SCAN_VARIANT >= 20 and plain else "bimamba_scan_fwd_kernel")
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "6.65 TB/s (of fallback)",
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_scan_ms,
                "launches_timed": len(scan_ms), "share_of_step": avg_scan_ms * len(scan_ms) / ms_dev if scan_ms else None}
        # the pipe that actually binds the scan: MUFU ex2, one per (token, channel, state)
        try:
            mufu = CF.microbench(0)
            ex2_per_launch = E * N * 2 * calls_per_launch * a.batch * a.seqlen
            roof["mufu"] = {"ex2_per_s_measured_peak": mufu, "ffma_per_s_measured_peak": CF.microbench(1),
                            "achieved_scan_ex2_per_s": ex2_per_launch / (avg_scan_ms * 1e-3) if scan_ms else None}
            if scan_ms:
                roof["mufu"]["frac"] = roof["mufu"]["achieved_scan_ex2_per_s"] / mufu
        except Exception as exc:  # noqa: BLE001
            roof["mufu"] = {"error": str(exc)}

        cpu = None
        if not a.no_cpu_baseline:
            rate, cores, sample = cpu_scan_rate(12.0, L=4096, min_calls=2)
            cpu = {"value": rate / scans_per_nt(a), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": sample + f"; nt/s = token-directions/s / {scans_per_nt(a)} (scan only)",
                   "token_directions_per_s": rate}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True,
            "scaling": "strong" if shard_seq else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "mode": a.mode,
            "config": {"workload": workload_name(a), "parallelism": (f"sp{world} (one sequence sharded on the sequence axis, 2 tiny all_gathers per layer)"
                                       if shard_seq else f"dp{world} (independent sequences per GPU)"),
                       "l2": "per-step working set > 1 GB >> 126 MB L2; 4 rotating input batches",
                       "launch": "CUDA graph replay" if graphed else "eager",
                       "scan_variant": 4 if scan_v4 else (CF.SCAN_VARIANT or 3)},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": a.batch * a.seqlen * 8 * world,
                    "d2h_bytes_per_step": (4 if train else a.batch * a.seqlen * cfg.vocab_size * 4) * world},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": sampler.summary() if sampler else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
