"""world_size-2 (and 3) gloo tests of the sequence-sharding host logic (caduceus_b200/seqshard.py) on CPU:
halo exchange + zero-carry pass + all_gather + carry composition + second pass must reproduce the unsharded scan.
The per-shard scan itself is stood in by a small torch recurrence with the C-ABI kernel's contract
(halo, h0 in; hlast, sum(dt) out) — the real kernel is covered by tests/test_gpu_parity.py and, across two GPUs,
by tests/test_gpu_multi.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from caduceus_b200 import seqshard

LOG2E = 1.4426950408889634


def local_scan(x, z, dt_raw, Bm, Cm, p, rev, halo=None, h0=None):
    """One job on one shard, logical-time semantics of cad_bimamba_scan_fwd (include/caduceus_b200.h).
    x, z, dt_raw: (E, L); Bm, Cm: (N, L).  Returns out (E, L), hlast (E, N), dtsum (E)."""
    E, L = x.shape
    if rev:
        x, z, dt_raw, Bm, Cm = (t.flip(-1) for t in (x, z, dt_raw, Bm, Cm))
    pre = torch.zeros(E, 3) if halo is None else halo
    xp = torch.cat([pre, x], dim=1)
    c = p["conv_b"][:, None] + sum(p["conv_w"][:, k:k + 1] * xp[:, k:k + L] for k in range(4))
    u = F.silu(c)
    dt = F.softplus(dt_raw + p["dt_b"][:, None])
    h = torch.zeros(E, p["A2"].shape[1]) if h0 is None else h0.clone()
    ys = []
    for t in range(L):
        h = torch.exp2(dt[:, t:t + 1] * p["A2"]) * h + dt[:, t:t + 1] * u[:, t:t + 1] * Bm[None, :, t]
        ys.append((h * Cm[None, :, t]).sum(-1) + p["D"] * u[:, t])
    out = torch.stack(ys, dim=1) * F.silu(z)
    if rev:
        out = out.flip(-1)
    return out, h, dt.sum(-1)


def make_problem(L, E=12, N=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = dict(conv_w=0.5 * torch.randn(E, 4, generator=g), conv_b=0.1 * torch.randn(E, generator=g),
             dt_b=torch.randn(E, generator=g) - 2.0, A2=-(torch.rand(E, N, generator=g) * 2 + 0.05) * LOG2E,
             D=torch.randn(E, generator=g))
    data = dict(x=torch.randn(1, E, L, generator=g), z=torch.randn(1, E, L, generator=g),
                dt_raw=torch.randn(2, E, L, generator=g), B=torch.randn(2, N, L, generator=g),
                C=torch.randn(2, N, L, generator=g))
    return p, data


def _worker(rank, world, port, L, out_q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, data = make_problem(L)
        seq_of_job = torch.tensor([0, 0], dtype=torch.int32)
        rev_of_job = torch.tensor([0, 1], dtype=torch.int32)       # job 0 left-to-right, job 1 right-to-left
        pset = torch.tensor([0, 0], dtype=torch.int32)
        Ls = L // world
        sl = slice(rank * Ls, (rank + 1) * Ls)
        with seqshard.sequence_parallel() as ctx:
            assert seqshard.current() is ctx and ctx.rank == rank and ctx.world == world
            halo = seqshard.gather_halo(data["x"][..., sl], Ls, seq_of_job, rev_of_job, ctx)
            hl, ds = [], []
            for j in range(2):
                _, h, s = local_scan(data["x"][0, :, sl], data["z"][0, :, sl], data["dt_raw"][j, :, sl],
                                     data["B"][j, :, sl], data["C"][j, :, sl], p, bool(rev_of_job[j]), halo=halo[j])
                hl.append(h); ds.append(s)
            h0 = seqshard.gather_carry(torch.stack(hl), torch.stack(ds), p["A2"][None], pset, rev_of_job, ctx)
            outs = []
            for j in range(2):
                o, _, _ = local_scan(data["x"][0, :, sl], data["z"][0, :, sl], data["dt_raw"][j, :, sl],
                                     data["B"][j, :, sl], data["C"][j, :, sl], p, bool(rev_of_job[j]), halo=halo[j],
                                     h0=h0[j])
                outs.append(o)
        assert seqshard.current() is None
        gathered = [torch.empty_like(torch.stack(outs)) for _ in range(world)]
        dist.all_gather(gathered, torch.stack(outs))
        if rank == 0:
            out_q.put(torch.cat(gathered, dim=-1))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_matches_unsharded(world):
    L = 48 * world
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get()
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p, data = make_problem(L)
    ref = torch.stack([local_scan(data["x"][0], data["z"][0], data["dt_raw"][j], data["B"][j], data["C"][j], p, bool(j))[0]
                       for j in range(2)])
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5), (got - ref).abs().max()


def test_compose_carry_is_the_affine_composition():
    torch.manual_seed(0)
    world, njobs, E, N = 4, 2, 3, 5
    H = torch.randn(world, njobs, E, N)
    S = torch.rand(world, njobs, E)
    A2 = -torch.rand(njobs, E, N)
    rev = torch.tensor([0, 1])
    for rank in range(world):
        h0 = seqshard.compose_carry(H, S, A2, rev, rank)
        f = torch.zeros(E, N)
        for j in range(rank):
            f = torch.exp2(A2[0] * S[j, 0][:, None]) * f + H[j, 0]
        b = torch.zeros(E, N)
        for j in range(world - 1, rank, -1):
            b = torch.exp2(A2[1] * S[j, 1][:, None]) * b + H[j, 1]
        assert torch.allclose(h0[0], f) and torch.allclose(h0[1], b)


# ---- backward: adjoint carry + halo gradient exchange (sequence-sharded TRAINING) ------------------------------------
def _bwd_worker(rank, world, port, L, out_q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, data = make_problem(L)
        dout = torch.randn(2, p["conv_b"].shape[0], L, generator=torch.Generator().manual_seed(5))
        seq_of_job = torch.tensor([0, 0], dtype=torch.int32)
        rev_of_job = torch.tensor([0, 1], dtype=torch.int32)
        pset = torch.tensor([0, 0], dtype=torch.int32)
        Ls = L // world
        sl = slice(rank * Ls, (rank + 1) * Ls)
        pl = {k: v.clone().requires_grad_() for k, v in p.items()}
        xj = [data["x"][0, :, sl].clone().requires_grad_() for _ in range(2)]       # per-job copies -> per-job dx
        z = data["z"][0, :, sl].clone().requires_grad_()
        dtr, Bm, Cm = (data[k][:, :, sl].clone().requires_grad_() for k in ("dt_raw", "B", "C"))
        with seqshard.sequence_parallel() as ctx:
            # forward, the kernel contract of _BiMambaCoreFn.forward: halo, zero-carry state pass, gather, true-carry scan
            halo = seqshard.gather_halo(data["x"][..., sl], Ls, seq_of_job, rev_of_job, ctx).requires_grad_()
            with torch.no_grad():
                st = [local_scan(xj[j], z, dtr[j], Bm[j], Cm[j], p, bool(rev_of_job[j]), halo=halo[j]) for j in range(2)]
            h0, dt_all = seqshard.gather_carry(torch.stack([s[1] for s in st]), torch.stack([s[2] for s in st]),
                                               p["A2"][None], pset, rev_of_job, ctx, return_dtsum=True)
            h0 = h0.requires_grad_()
            res = [local_scan(xj[j], z, dtr[j], Bm[j], Cm[j], pl, bool(rev_of_job[j]), halo=halo[j], h0=h0[j])
                   for j in range(2)]
            obj = sum((res[j][0] * dout[j, :, sl]).sum() for j in range(2))
            # backward: Dh (what cad_bimamba_scan_adjoint computes) -> gather/compose -> backward with dhlast
            dh = torch.autograd.grad(obj, h0, retain_graph=True)[0]
            dhlast = seqshard.gather_adjoint(dh, dt_all, p["A2"][None], pset, rev_of_job, ctx)
            (obj + sum((res[j][1] * dhlast[j]).sum() for j in range(2))).backward()
            dx = torch.stack([t.grad for t in xj])
            seqshard.exchange_halo_grad(halo.grad, dx, Ls, seq_of_job, rev_of_job, ctx)
            lin = torch.nn.Module()
            for k, v in pl.items():
                lin.register_parameter(k, torch.nn.Parameter(v.detach().clone()))
                getattr(lin, k).grad = v.grad.clone()
            seqshard.all_reduce_grads(lin, ctx)
        tok = torch.cat([dx.sum(0)[None], z.grad[None], dtr.grad], 0)                                        # (4, E, Ls)
        bcg = torch.cat([Bm.grad, Cm.grad], 0)                                                             # (4, N, Ls)
        g1 = [torch.empty_like(tok) for _ in range(world)]
        g2 = [torch.empty_like(bcg) for _ in range(world)]
        dist.all_gather(g1, tok)
        dist.all_gather(g2, bcg)
        if rank == 0:
            out_q.put((torch.cat(g1, -1), torch.cat(g2, -1), {k: getattr(lin, k).grad for k in pl}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_backward_matches_unsharded(world):
    L = 48 * world
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_bwd_worker, args=(r, world, port, L, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    tok, bcg, pg = q.get()
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    p, data = make_problem(L)
    dout = torch.randn(2, p["conv_b"].shape[0], L, generator=torch.Generator().manual_seed(5))
    pl = {k: v.clone().requires_grad_() for k, v in p.items()}
    x, z = data["x"][0].clone().requires_grad_(), data["z"][0].clone().requires_grad_()
    dtr, Bm, Cm = (data[k].clone().requires_grad_() for k in ("dt_raw", "B", "C"))
    sum((local_scan(x, z, dtr[j], Bm[j], Cm[j], pl, bool(j))[0] * dout[j]).sum() for j in range(2)).backward()
    want_tok = torch.cat([x.grad[None], z.grad[None], dtr.grad], 0)
    want_bc = torch.cat([Bm.grad, Cm.grad], 0)
    assert torch.allclose(tok, want_tok, rtol=1e-4, atol=1e-5), (tok - want_tok).abs().max()
    assert torch.allclose(bcg, want_bc, rtol=1e-4, atol=1e-5), (bcg - want_bc).abs().max()
    for k in pl:
        assert torch.allclose(pg[k], pl[k].grad, rtol=1e-4, atol=1e-5), (k, (pg[k] - pl[k].grad).abs().max())


def test_conv_halo_grad_matches_autograd():
    """CF.conv_halo_grad (plain torch, used by the sharded backward) vs autograd through the conv + SiLU definition."""
    from caduceus_b200 import functional as CF
    g = torch.Generator().manual_seed(2)
    nseq, E, L, P = 2, 5, 11, 2
    xz = torch.randn(nseq, 2 * E, 16, generator=g)
    jobs = (torch.tensor([0, 0, 1, 1], dtype=torch.int32), torch.tensor([0, 1, 0, 1], dtype=torch.int32),
            torch.tensor([0, 1, 1, 0], dtype=torch.int32))
    w, b = torch.randn(P, E, 4, generator=g), torch.randn(P, E, generator=g)
    halo = torch.randn(4, E, 3, generator=g).requires_grad_()
    du = torch.randn(4, E, 16, generator=g)
    obj = 0.0
    for j in range(4):
        x = xz[jobs[0][j], :E, :L]
        du_j = du[j, :, :L]
        if jobs[2][j]:
            x, du_j = x.flip(-1), du_j.flip(-1)
        xp = torch.cat([halo[j], x], dim=1)
        c = b[jobs[1][j]][:, None] + sum(w[jobs[1][j]][:, k:k + 1] * xp[:, k:k + L] for k in range(4))
        obj = obj + (F.silu(c) * du_j).sum()
    obj.backward()
    got = CF.conv_halo_grad(xz, du, halo.detach(), w, b, jobs, L)
    assert torch.allclose(got, halo.grad, rtol=1e-5, atol=1e-6), (got - halo.grad).abs().max()
