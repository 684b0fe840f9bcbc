#!/usr/bin/env python
"""Headline benchmark: nucleotides/s of the Caduceus forward at seq_len=131072, d_model=256, n_layer=16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model ps|ph] [--seqlen L] [--impl reference]

Workload (BASELINE.json configs[2]): Caduceus-PS (rcps=true) d_model=256 n_layer=16 seq_len=131072 bf16
forward, batch 1 per GPU, random-init weights (reference init), synthetic hg38-shaped ids.
A "step" = one forward over one batch.  N > 1 (one process per GPU, torchrun): ONE sequence of --seqlen tokens per batch row is
sharded on the sequence axis over the N ranks (`--shard seq`, the default for N > 1; north_star / SURVEY.md §8e) -> "scaling":
"strong".  The per-layer exchanges (conv halo, boundary state) are stores into the peers' memory over NVLink by this library's own
kernels (csrc/peer_exchange.cu), and the whole sharded forward replays from ONE CUDA graph per rank; NCCL carries only the timing
barrier.  `--shard none` runs N independent replicas instead (the reference's own data-parallel mode, ref:train.py:629-639).

`--impl reference` times the reference's CPU path for the same metric: the oracle port of upstream `mamba_inner_ref`
(oracle/mamba_ssm/ops/selective_scan_interface.py: conv + SiLU, x_proj, dt_proj, `selective_scan_ref`, out_proj — the operator the
GPU arm fuses) at L = 4096 on all host threads, rank 0 only; the GPU arm's `cpu_baseline` leg times the SAME sample.
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
    ap.add_argument("--graph", action="store_true", help="replay the forward from a CUDA graph (default for the sharded forward)")
    ap.add_argument("--no-graph", action="store_true", help="eager launches even for the sharded forward")
    ap.add_argument("--scan-variant", type=int, default=None, choices=[0, 3, 20],
                    help="force the forward-scan kernel (cad_scan_fwd_args.variant); default: chosen per call (CF.choose_scan_variant)")
    ap.add_argument("--shard", default="auto", choices=["auto", "none", "seq"],
                    help="seq: ONE sequence of --seqlen sharded on the sequence axis over all ranks (strong scaling; auto = seq "
                         "when N > 1); none: independent replicas")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="sharded forward: halo / boundary-state exchange through NVLink peer memory (this library's kernels) or "
                         "NCCL all_gather (training always uses NCCL)")
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
CPU_SAMPLE_L = 4096


def cpu_inner_rate(seconds_budget, d_model=256, min_calls=1):
    """token-directions/s of the oracle's mamba_inner_ref (conv + SiLU -> x_proj -> dt_proj -> selective_scan_ref -> out_proj;
    fp32, all host threads) on ONE fixed sample: B = 1, L = 4096, E = 2 * d_model, N = 16.  Used by both arms."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mamba_ssm.ops.selective_scan_interface import mamba_inner_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L, E, N, D = CPU_SAMPLE_L, 2 * d_model, 16, d_model
    R = (D + 15) // 16
    g = torch.Generator().manual_seed(0)
    xz = torch.randn(1, 2 * E, L, generator=g)
    conv_w, conv_b = 0.5 * torch.randn(E, 1, 4, generator=g), 0.1 * torch.randn(E, generator=g)
    w_x, w_dt = torch.randn(R + 2 * N, E, generator=g) * E ** -0.5, torch.randn(E, R, generator=g) * R ** -0.5
    w_out = torch.randn(D, E, generator=g) * E ** -0.5
    A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(E, 1)
    Dp, dbias = torch.ones(E), torch.log(torch.expm1(torch.exp(torch.rand(E, generator=g) * 4.6 - 6.9)))
    calls, t_total = 0, 0.0
    while calls < min_calls or t_total < seconds_budget:
        t0 = time.perf_counter()
        mamba_inner_ref(xz, conv_w, conv_b, w_x, w_dt, w_out, None, A, None, None, Dp, delta_bias=dbias, delta_softplus=True)
        t_total += time.perf_counter() - t0
        calls += 1
    return calls * L / t_total, cores, (f"{calls} x mamba_inner_ref(B=1, D={D}, E={E}, N={N}, L={L}) fp32, {cores} threads "
                                        "(conv + x_proj + dt_proj + selective_scan_ref + out_proj)")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step, rates = [], []
    sample = ""
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        rate, cores, sample = cpu_inner_rate(0.0, a.d_model)
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
                   "same_config": False,
                   "note": "CPU oracle port of upstream mamba_inner_ref (one Mamba direction: conv, projections, scan; no norm / "
                           f"embedding / head) on a fixed L={CPU_SAMPLE_L} sample; nt/s = token-directions/s / {scans_per_nt(a)} "
                           "Mamba passes per nucleotide — context for the GPU number, not a like-for-like arm"},
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

    shard_seq = a.shard in ("seq", "auto") and world > 1
    peer = None
    if shard_seq:       # every rank holds tokens [rank*L/P, (rank+1)*L/P) of the SAME sequences
        from caduceus_b200 import seqshard
        assert a.seqlen % world == 0, "--seqlen must be a multiple of the number of ranks"
        Ls = a.seqlen // world
        if not train and a.exchange == "peer":
            # halo / boundary-state exchange through the peers' memory over NVLink (csrc/peer_exchange.cu): no collective on the
            # data path, and the whole sharded forward is capturable in one CUDA graph
            nstrand = 2 if a.model == "ps" else 1
            try:
                peer = seqshard.PeerExchange.create(nseq_max=a.batch * nstrand, njobs_max=a.batch * nstrand * 2, E=2 * a.d_model,
                                                    N=16, device=dev)
                ok = torch.ones(1, device=dev)
            except Exception as exc:  # noqa: BLE001  (no peer mapping on this box: say so and measure the NCCL exchange instead)
                print(f"[bench] rank {rank}: symmetric-memory workspace unavailable ({exc!r}); using the NCCL exchange", file=sys.stderr)
                peer, ok = None, torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks or none
            if ok.item() == 0:
                peer = None
        host_ids = [make_ids(torch, a.batch, a.seqlen, 100 + i)[:, rank * Ls:(rank + 1) * Ls].contiguous().pin_memory()
                    for i in range(nbuf)]
        dev_ids = [h.to(dev) for h in host_ids]
        host_out = torch.empty(a.batch, Ls, cfg.vocab_size, dtype=torch.float32).pin_memory()
        _model = model

        def model(ids):          # noqa: F811  (sequence-parallel wrapper around the same module)
            with seqshard.sequence_parallel(peer=peer):
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

    # the device -> host read of step i's logits runs on its own stream while step i + 1 computes (the copy engine is idle
    # otherwise); the timed region ends only when the last read has landed (timed() joins the stream before its end event)
    d2h_stream = torch.cuda.Stream()

    def step_e2e(i):
        ids = host_ids[i % nbuf].to(dev, non_blocking=True)
        if train:
            loss = train_step(ids)
            host_loss.copy_(loss.detach(), non_blocking=True)
            return loss
        with torch.no_grad():
            logits = model(ids).logits
        d2h_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(d2h_stream):
            host_out.copy_(logits, non_blocking=True)
        logits.record_stream(d2h_stream)
        return logits

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        torch.cuda.current_stream().wait_stream(d2h_stream)
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
    # CUDA-graph replay: on request for the single-GPU / replica forward, by default for the peer-exchange sharded forward (its
    # per-rank GPU work per layer is a fraction of a millisecond: eager launches would bound the step by the host)
    if not train and not a.no_graph and (a.graph or peer is not None):
        from caduceus_b200.graphs import GraphedForward
        CF.LAUNCHES = 0
        step_device(0)
        launches_per_step = CF.LAUNCHES
        try:
            gfwd = GraphedForward(model, dev_ids[0])
            graphed = True
        except Exception as exc:  # noqa: BLE001
            print(f"[bench] rank {rank}: CUDA-graph capture failed ({exc!r}); eager launches", file=sys.stderr)
        if graphed:
            eager_device = step_device

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
    CF.XPROJ_EVENTS = None if (graphed or train) else []
    ms_dev = timed(step_device, a.steps)
    launches = launches_per_step * a.steps if graphed else CF.LAUNCHES     # a replayed graph re-runs the captured launches
    scan_events, CF.SCAN_EVENTS = (CF.SCAN_EVENTS or []), None
    xproj_events, CF.XPROJ_EVENTS = (CF.XPROJ_EVENTS or []), None
    ms_e2e = timed(step_e2e, a.steps)
    scan_timed_in = "the timed region"
    if graphed:
        # events cannot be recorded inside a replayed graph: time the scan launches of two EAGER steps after the timed regions
        # (same kernels, same shapes; every rank runs them — the sharded forward exchanges with its peers)
        CF.SCAN_EVENTS = []
        CF.XPROJ_EVENTS = None if train else []
        barrier()
        for i in range(2):
            eager_device(i)
        torch.cuda.synchronize()
        scan_events, CF.SCAN_EVENTS = CF.SCAN_EVENTS, None
        xproj_events, CF.XPROJ_EVENTS = (CF.XPROJ_EVENTS or []), None
        scan_timed_in = "2 eager steps after the timed region (the timed steps replay a CUDA graph)"
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
        L_launch = a.seqlen // world if shard_seq else a.seqlen                # tokens one launch of THIS rank covers
        bytes_per_launch = 2 * (8 * E + 4 * N) * calls_per_launch * a.batch * L_launch
        scan_ms = [s.elapsed_time(e) for s, e in scan_events]
        avg_scan_ms = sum(scan_ms) / max(len(scan_ms), 1)
        nstrand = 2 if a.model == "ps" else 1
        v_run = 3 if train else CF.choose_scan_variant(torch.bfloat16, N, a.batch * nstrand * 2, E, L_launch)
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
                # measured traffic exists per kernel: no entry for the running kernel -> null, not another kernel's number
                traffic = tj.get(a.model + ("" if v_run == 3 else f"_v{v_run}"), {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        kname = ("bimamba_scan_fwd_v20_kernel + seg_carry_kernel + scan_fixup_kernel (lane = channel: the whole segmented scan is timed)"
                 if v_run == 20 else "bimamba_scan_fwd_kernel (one channel per warp, time-parallel)")
        roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if achieved else None, "traffic": traffic,
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "6.65 TB/s (of fallback)",
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_scan_ms,
                "launches_timed": len(scan_ms), "timed_in": scan_timed_in,
                "share_of_step": (avg_scan_ms * a.n_layer) / (ms_dev / a.steps) if scan_ms else None}
        # second kernel of the step: conv + SiLU -> x_proj -> dt_proj on tcgen05 / TMEM (csrc/xproj.cu), HBM-bound.  Per job and
        # token it reads x (2E bytes), writes delta (2E bytes) and the fp32 B / C rows state-major (8N bytes) and, for the lane =
        # channel scan, token-major as well (8N bytes)
        xp_ms = [s.elapsed_time(e) for s, e in xproj_events]
        if xp_ms:
            njobs_launch = a.batch * nstrand * 2
            xp_bytes = njobs_launch * L_launch * (4 * E + 8 * N * (2 if v_run == 20 else 1))
            avg_xp = sum(xp_ms) / len(xp_ms)
            roof["second_kernel"] = {"kernel": "conv_xproj_umma_kernel (tcgen05.mma + TMEM accumulators, TMA in / out)", "bound": "hbm",
                                     "achieved": xp_bytes / (avg_xp * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": xp_bytes / (avg_xp * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": xp_bytes,
                                     "avg_launch_ms": avg_xp, "launches_timed": len(xp_ms),
                                     "share_of_step": (avg_xp * a.n_layer) / (ms_dev / a.steps)}
        # the pipe that actually binds the scan: MUFU ex2, one per (token, channel, state)
        try:
            mufu = CF.microbench(0)
            ex2_per_launch = E * N * 2 * calls_per_launch * a.batch * L_launch
            roof["mufu"] = {"ex2_per_s_measured_peak": mufu, "ffma_per_s_measured_peak": CF.microbench(1),
                            "achieved_scan_ex2_per_s": ex2_per_launch / (avg_scan_ms * 1e-3) if scan_ms else None}
            if scan_ms:
                roof["mufu"]["frac"] = roof["mufu"]["achieved_scan_ex2_per_s"] / mufu
        except Exception as exc:  # noqa: BLE001
            roof["mufu"] = {"error": str(exc)}

        cpu = None
        if not a.no_cpu_baseline:
            rate, cores, sample = cpu_inner_rate(12.0, a.d_model, min_calls=2)
            cpu = {"value": rate / scans_per_nt(a), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": sample + f"; nt/s = token-directions/s / {scans_per_nt(a)} Mamba passes per nucleotide",
                   "token_directions_per_s": rate}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True,
            "scaling": "strong" if shard_seq else "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "mode": a.mode,
            "config": {"workload": workload_name(a),
                       "parallelism": ((f"sp{world}: one sequence per batch row sharded on the sequence axis; per layer one conv-halo and "
                                        "one boundary-state exchange, " +
                                        ("stores into the peers' memory over NVLink by the library's own kernels, no collective"
                                         if peer is not None else "NCCL all_gather"))
                                       if shard_seq else f"dp{world} (independent sequences per GPU)"),
                       "l2": "per-step working set > 1 GB >> 126 MB L2; 4 rotating input batches",
                       "launch": "CUDA graph replay" if graphed else "eager",
                       "scan_variant": v_run},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": a.batch * a.seqlen * 8 * (1 if shard_seq else world),
                    "d2h_bytes_per_step": (4 * world if train else a.batch * a.seqlen * cfg.vocab_size * 4 * (1 if shard_seq else world)),
                    "copies": ("H2D of the step's ids from pinned memory on the compute stream; D2H of its result " +
                               ("on the compute stream" if (graphed or train) else
                                "on a second stream, overlapped with the next step's compute; the timed region ends after the last copy has landed"))},
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
