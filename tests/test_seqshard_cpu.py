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
