"""Sequence-sharded TRAINING (SURVEY.md §8e "Backward"), -m gpu.

Kernel level (one process): two shards of one sequence, carried by hand — cad_bimamba_scan_bwd with (h0, dhlast) on the
shards must reproduce the unsharded backward, and cad_bimamba_scan_adjoint must equal the dh0 the backward exports.
Model level: `world` processes (NCCL when there are enough GPUs, otherwise gloo with the ranks sharing cuda:0 — the
collectives are host-staged then, the kernels are the same) train one step on their shard; loss and ALL parameter
gradients must match the single-process unsharded step."""
import os
import socket

import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def _close(got, ref, rtol, atol, what):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().clamp_min(1e-30)
    bound = atol * scale + rtol * ref.abs()
    assert torch.all(err <= bound), f"{what}: max err {err.max():.3e} (ref max {scale:.3e})"


def _problem(L, E, dtype, seed=0):
    from caduceus_b200 import functional as CF
    g = torch.Generator().manual_seed(seed)
    N, P = 16, 2
    Lp, Lbc = CF.round_up(L, 16), CF.round_up(L, 32)
    xz = torch.zeros(1, 2 * E, Lp)
    xz[..., :L] = torch.randn(1, 2 * E, L, generator=g)
    delta = torch.zeros(2, E, Lp)
    delta[..., :L] = 0.5 * torch.randn(2, E, L, generator=g)
    bc = torch.zeros(2, 2 * N, Lbc)
    bc[..., :L] = torch.randn(2, 2 * N, L, generator=g)
    dout = torch.zeros(2, E, Lp)
    dout[..., :L] = torch.randn(2, E, L, generator=g)
    packed = (0.5 * torch.randn(P, E, 4, generator=g), 0.1 * torch.randn(P, E, generator=g),
              torch.randn(P, E, generator=g) - 3.0, -(torch.rand(P, E, N, generator=g) * 4 + 0.02) * LOG2E,
              torch.randn(P, E, generator=g))
    jobs = tuple(torch.tensor(v, dtype=torch.int32, device=DEV) for v in ([0, 0], [0, 1], [0, 1]))
    dev = lambda t, dt=None: t.to(DEV, dt) if dt else t.to(DEV)       # noqa: E731
    return (dev(xz, dtype), dev(delta, dtype), dev(bc), dev(dout, dtype), tuple(dev(t) for t in packed), jobs)


def _cut(t, lo, hi, mult):
    """Contiguous shard copy [lo, hi) of the last axis, zero padded to a multiple of `mult`."""
    from caduceus_b200 import functional as CF
    n = hi - lo
    out = torch.zeros(*t.shape[:-1], CF.round_up(n, mult), device=t.device, dtype=t.dtype)
    out[..., :n] = t[..., lo:hi]
    return out


@pytest.mark.parametrize("dtype,L,La", [(torch.float32, 1300, 640), (torch.float32, 96, 48), (torch.bfloat16, 1300, 640)])
def test_two_shards_with_h0_and_dhlast_reproduce_the_unsharded_backward(dtype, L, La):
    from caduceus_b200 import functional as CF
    E = 24
    xz, delta, bc, dout, packed, jobs = _problem(L, E, dtype)
    out_f, _, _, cs_f = CF.scan_fwd(xz, delta, bc, packed, jobs, L, want_chunk_state=True)
    full = CF.scan_bwd(xz, delta, bc, dout, packed, jobs, L, cs_f)

    bounds = [(0, La), (La, L)]
    sh = [dict(xz=_cut(xz, lo, hi, 16), delta=_cut(delta, lo, hi, 16), bc=_cut(bc, lo, hi, 32), dout=_cut(dout, lo, hi, 16),
               L=hi - lo) for lo, hi in bounds]
    x = xz[0, :E]
    zeros3 = torch.zeros(E, 3, device=DEV, dtype=dtype)
    # job 0 runs left to right (shard 1 follows shard 0), job 1 right to left (shard 0 follows shard 1)
    sh[0]["halo"] = torch.stack([zeros3, x[:, La:La + 3].flip(-1)]).contiguous()
    sh[1]["halo"] = torch.stack([x[:, La - 3:La], zeros3]).contiguous()
    for s in sh:
        _, s["hl"], s["ds"], _ = CF.scan_fwd(s["xz"], s["delta"], s["bc"], packed, jobs, s["L"], halo=s["halo"],
                                             state_only=True)
    z = torch.zeros_like(sh[0]["hl"][0])
    sh[0]["h0"] = torch.stack([z, sh[1]["hl"][1]])
    sh[1]["h0"] = torch.stack([sh[0]["hl"][0], z])
    for s in sh:
        s["out"], _, _, s["cs"] = CF.scan_fwd(s["xz"], s["delta"], s["bc"], packed, jobs, s["L"], halo=s["halo"],
                                              h0=s["h0"], want_chunk_state=True)
        s["dh"] = CF.scan_adjoint(s["xz"], s["delta"], s["bc"], s["dout"], packed, jobs, s["L"])
        # the adjoint kernel == what the backward itself exports as dh0 when nothing enters from the next shard
        g0 = CF.scan_bwd(s["xz"], s["delta"], s["bc"], s["dout"], packed, jobs, s["L"], s["cs"], halo=s["halo"],
                         h0=s["h0"], want_dh0=True)
        _close(s["dh"], g0[7], 1e-3, 1e-5, "adjoint kernel vs backward dh0")
    sh[0]["dhlast"] = torch.stack([sh[1]["dh"][0], z])
    sh[1]["dhlast"] = torch.stack([z, sh[0]["dh"][1]])
    for s in sh:
        s["g"] = CF.scan_bwd(s["xz"], s["delta"], s["bc"], s["dout"], packed, jobs, s["L"], s["cs"], halo=s["halo"],
                             h0=s["h0"], dhlast=s["dhlast"])
    rt, at = (2e-3, 2e-5) if dtype == torch.float32 else (3e-2, 2e-2)
    cat = lambda key_or_idx, src: torch.cat([src(s)[..., :s["L"]] for s in sh], dim=-1)      # noqa: E731
    _close(cat(None, lambda s: s["out"]), out_f[..., :L], rt, at, "forward with carried h0")
    for i, name in enumerate(["dz", "du", "ddelta", "dbc"]):
        _close(cat(i, lambda s: s["g"][i]), full[i][..., :L], rt, at, name)
    for i, name in [(4, "ddt_b"), (5, "dA2"), (6, "dD")]:
        _close(sh[0]["g"][i] + sh[1]["g"][i], full[i], rt, 10 * at, name)


# ---- model level -------------------------------------------------------------------------------------------------------
def _train_worker(rank, world, port, tag, backend, q):
    import torch.distributed as dist
    import torch.nn.functional as F
    import caduceus
    from caduceus_b200 import seqshard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fx = golden(f"model_{tag}.pt")
        cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
        model = caduceus.CaduceusForMaskedLM(cfg)
        model.load_state_dict(fx["state_dict"])
        model = model.to(dev).train()
        g = torch.Generator().manual_seed(0)
        Ls = 600                                     # not a multiple of the 512-token chunk
        L = Ls * world
        ids = torch.randint(7, 11, (2, L), generator=g)
        labels = ids.clone()
        labels[torch.rand(ids.shape, generator=g) > 0.3] = 4
        nvalid = (labels != 4).sum().item()
        sl = slice(rank * Ls, (rank + 1) * Ls)
        with seqshard.sequence_parallel() as ctx:
            logits = model(ids[:, sl].to(dev)).logits
            loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), labels[:, sl].reshape(-1).to(dev),
                                   ignore_index=4, reduction="sum") / nvalid
            loss.backward()
            seqshard.all_reduce_grads(model, ctx)
        total = loss.detach().clone()
        if backend == "nccl":
            dist.all_reduce(total)
        else:
            t = total.cpu()
            dist.all_reduce(t)
            total = t
        if rank == 0:
            # numpy payloads: torch tensors would travel by file descriptor and need this process alive at receive time
            sharded = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
            model.zero_grad(set_to_none=True)
            logits = model(ids.to(dev)).logits
            ref = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), labels.reshape(-1).to(dev), ignore_index=4)
            ref.backward()
            full = {n: p.grad.detach().float().cpu().numpy() for n, p in model.named_parameters() if p.grad is not None}
            q.put((total.item(), ref.item(), sharded, full))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tag,world", [("ph_small", 2), ("ps_small", 3)])
def test_sequence_sharded_train_step_matches_unsharded(tag, world):
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, tag, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        loss_s, loss_f, sharded, full = q.get(timeout=300)
    finally:
        for p in procs:
            p.join(180)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert abs(loss_s - loss_f) < 1e-4 * max(1.0, abs(loss_f)), (loss_s, loss_f)
    assert set(sharded) == set(full) and len(full) >= 10
    for n in full:
        _close(torch.from_numpy(sharded[n]), torch.from_numpy(full[n]), 2e-2, 2e-3, f"d {n}")
