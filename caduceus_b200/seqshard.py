"""Sequence-axis sharding of the BiMamba path across ranks (one process per GPU, NCCL).

The reference itself only scales by data parallelism (ref:train.py:629-639); sharding ONE long sequence over the
8 GPUs of a box is what BASELINE.json's north_star adds.  Rank k owns the physical tokens [k*Ls, (k+1)*Ls) of every
sequence, all channels; weights are replicated.  Everything on the path is token-local except, per BiMamba call,

  (i)  the conv halo: the 3 `x` samples that logically precede the shard, and
  (ii) the scan carry: the (E x N) state at the shard boundary of every (sequence, direction) job.

A literal "send h_end to the next rank" chain serialises the ranks (rank k cannot finish before rank k-1).  The
recurrence is affine in the carried state, so instead every rank
   1. scans its shard from a ZERO state, producing the outputs of that zero-carry scan plus the end state H_k and
      sum(dt)_k, from which the shard's total decay is P_k = exp2(A2 * sum(dt)_k);
   2. takes part in ONE small all_gather of (H, sum dt) — 2 x (32 KiB + 2 KiB) per job —
   3. composes its true carry-in  h0_k = sum_{j<k} (prod_{j<i<k} P_i) H_j  locally (reversed order for reversed jobs),
   4. adds the contribution of h0_k to its outputs in place:  silu(z) * sum_n C * exp2(A2 * cumsum(dt)) * h0
      (csrc/scan_fixup.cu) — a term that only decays along the shard, so it is cut off per (channel, state) once it is
      below 2^-40 and typically touches a few percent of the shard.
All scans are shard-local and start together, so the P ranks run concurrently: time per layer ~ T_scan(L / P) * (1 + eps).
Measured on 2 x B200 (L = 131072, D = 256, 16 layers, bf16 forward): PS 50.8 -> 32.7 ms, Ph 37.6 -> 23.2 ms.

`sequence_parallel(group)` is the user-facing switch: inside it `bimamba_inner` (hence every Caduceus model of this
package) treats its input as the local shard.  Training works the same way (see the "backward" block below): the
forward keeps the TRUE carry (zero-carry state pass -> all_gather -> full scan from h0), the backward gathers the
per-shard adjoints once and runs the ordinary backward kernel with (h0, dhlast); parameter gradients are then summed
over the ranks with `all_reduce_grads`.
"""
import contextlib
import threading

import torch
import torch.distributed as dist

_TLS = threading.local()        # the active ShardContext is per THREAD (several virtual ranks may live in one test process)


class ShardContext:
    def __init__(self, group=None, peer=None):
        self.group = group
        self.rank = dist.get_rank(group) if peer is None else peer.rank
        self.world = dist.get_world_size(group) if peer is None else peer.world
        # gloo (CPU tests, or several ranks sharing one GPU in CI) moves CUDA tensors through host memory; the
        # production backend is NCCL, which takes the device buffers as they are
        self.host_staged = peer is None and dist.get_backend(group) == "gloo"
        # PeerExchange: the forward's two exchanges go through the peers' memory (csrc/peer_exchange.cu) instead of a collective
        self.peer = peer


class PeerExchange:
    """Symmetric workspace + peer address table for csrc/peer_exchange.cu (inference forward of a sequence-sharded model).

    `PeerExchange.create(group, ...)` allocates the workspace with torch's symmetric memory (CUDA VMM handles exchanged over the
    process group's store; every rank of one NVLink / NVSwitch box maps every peer's buffer), zero-fills it and barriers once.
    `PeerExchange.from_buffers(rank, buffers, ...)` builds the same object from explicit device buffers — used by the tests to
    run several virtual ranks inside one process on one GPU.  The kernels only need the table of base addresses."""

    def __init__(self, rank, world, ptrs, device, nseq_max, njobs_max, E, N, keep=()):
        from . import _lib
        self.rank, self.world = int(rank), int(world)
        self.table = torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=device)      # uint64 addresses
        self.ctx = _lib.PeerCtx(self.table.data_ptr(), self.rank, self.world, nseq_max, njobs_max, E, N)
        self.geometry = (nseq_max, njobs_max, E, N)
        self._keep = keep            # whatever owns the memory (symmetric-memory handle / buffers)

    @classmethod
    def create(cls, group=None, *, nseq_max, njobs_max, E, N, device=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import functional as CF
        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        nbytes = CF.peer_ws_bytes(world, nseq_max, njobs_max, E, N)
        buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        hdl = symm_mem.rendezvous(buf, group)
        buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)                       # nobody pushes before every workspace is zeroed
        return cls(rank, world, list(hdl.buffer_ptrs), device, nseq_max, njobs_max, E, N, keep=(buf, hdl))

    @classmethod
    def from_buffers(cls, rank, buffers, *, nseq_max, njobs_max, E, N):
        return cls(rank, len(buffers), [b.data_ptr() for b in buffers], buffers[0].device, nseq_max, njobs_max, E, N,
                   keep=tuple(buffers))

    def covers(self, nseq, njobs, E, N):
        g = self.geometry
        return nseq <= g[0] and njobs <= g[1] and E == g[2] and N == g[3]


def current():
    return getattr(_TLS, "ctx", None)


@contextlib.contextmanager
def sequence_parallel(group=None, peer=None):
    """Treat the sequence axis of every BiMamba call inside as sharded over `group` (default: WORLD).
    peer: a PeerExchange — the inference forward then exchanges halo and boundary states through the peers' memory
    (no collective, CUDA-graph capturable); training and fp32 paths keep the NCCL exchange."""
    prev = current()
    _TLS.ctx = ShardContext(group, peer)
    try:
        yield _TLS.ctx
    finally:
        _TLS.ctx = prev


def _all_gather(t, ctx):
    """(world, *t.shape) — one collective, one output buffer."""
    t = t.contiguous()
    if ctx.host_staged and t.is_cuda:
        return _all_gather(t.cpu(), ctx).to(t.device)
    out = torch.empty((ctx.world,) + tuple(t.shape), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out.view(ctx.world * t.shape[0], *t.shape[1:]) if t.dim() > 0 else out, t,
                                group=ctx.group)
    return out


_HALO_INDEX = {}


def _halo_index(seq_l, rev_l, rank, world, device):
    """Per job: which (rank, edge, sequence) supplies its halo, whether it is reversed, and whether it exists."""
    key = (seq_l, rev_l, rank, world, str(device))
    hit = _HALO_INDEX.get(key)
    if hit is None:
        src_rank = [rank + 1 if r else rank - 1 for r in rev_l]
        valid = [0 <= q < world for q in src_rank]
        src_rank = [q if ok else rank for q, ok in zip(src_rank, valid)]
        edge = [0 if r else 1 for r in rev_l]                      # reversed jobs take the successor's FIRST three
        mk = lambda v, dt: torch.tensor(v, dtype=dt, device=device)     # noqa: E731
        hit = (mk(src_rank, torch.long), mk(edge, torch.long), mk(list(seq_l), torch.long),
               mk(list(rev_l), torch.bool)[:, None, None], mk(valid, torch.bool)[:, None, None])
        _HALO_INDEX[key] = hit
    return hit


def gather_halo(x_rows, L, seq_of_job, rev_of_job, ctx):
    """x_rows (nseq, E, >=L): the conv input rows of this shard.  Returns halo (njobs, E, 3) = the 3 samples
    logically preceding the shard for each job (zeros at the ends of the full sequence), in logical order."""
    assert L >= 3, "sequence shards must hold at least 3 tokens"
    from . import functional as CF
    host = CF.JOB_HOST.get(seq_of_job.data_ptr())
    seq_l, rev_l = (host[0], host[2]) if host is not None else (tuple(seq_of_job.tolist()), tuple(rev_of_job.tolist()))
    edges = torch.stack([x_rows[..., 0:3], x_rows[..., L - 3:L]])              # (2, nseq, E, 3): first3, last3
    allv = _all_gather(edges, ctx)                                              # (world, 2, nseq, E, 3)
    src_rank, edge, seq_i, is_rev, valid = _halo_index(seq_l, rev_l, ctx.rank, ctx.world, x_rows.device)
    h = allv[src_rank, edge, seq_i]                                             # (njobs, E, 3)
    # left-to-right jobs: predecessor's LAST three, already in logical order; right-to-left: successor's FIRST three,
    # logical order = physical order reversed
    h = torch.where(is_rev, h.flip(-1), h)
    return torch.where(valid, h, torch.zeros_like(h)).contiguous()


def compose_carry(h_all, dtsum_all, A2_job, rev_of_job, rank):
    """h_all (world, njobs, E, N) zero-carry end states, dtsum_all (world, njobs, E), A2_job (njobs, E, N).
    Carry-in of `rank`:  h0 = sum over logical predecessors j of  exp2(A2 * sum(dt over the ranks between j and rank)) * H_j
    (the affine composition h <- exp2(A2 * sum dt_j) * h + H_j unrolled), vectorised over the ranks."""
    world = h_all.shape[0]
    idx = torch.arange(world, device=h_all.device)
    csum = torch.cumsum(dtsum_all, dim=0)                                        # inclusive over ranks
    total = csum[-1]
    # left-to-right jobs: predecessors j < rank, decay over ranks j+1 .. rank-1  ->  S(rank-1) - S(j)
    upto = csum[rank - 1] if rank > 0 else torch.zeros_like(total)
    between_f = (upto[None] - csum).clamp_min(0.0)
    w_f = torch.exp2(A2_job[None] * between_f[..., None]) * (idx < rank)[:, None, None, None]
    # right-to-left jobs: predecessors j > rank, decay over ranks rank+1 .. j-1  ->  S(j-1) - S(rank)
    csum_prev = csum - dtsum_all
    between_r = (csum_prev - csum[rank][None]).clamp_min(0.0)
    w_r = torch.exp2(A2_job[None] * between_r[..., None]) * (idx > rank)[:, None, None, None]
    rev = rev_of_job.to(torch.bool)[None, :, None, None]
    return (torch.where(rev, w_r, w_f) * h_all).sum(0).contiguous()


def gather_carry(hlast, dtsum, A2, pset_of_job, rev_of_job, ctx, return_dtsum=False):
    """All-gather (H, sum dt) of the zero-carry pass and compose this rank's carry-in state (njobs, E, N)."""
    E, N = hlast.shape[1], hlast.shape[2]
    packed = torch.cat([hlast.reshape(hlast.shape[0], -1), dtsum], dim=1)       # one message per rank
    allv = _all_gather(packed, ctx)
    h_all = allv[:, :, :E * N].reshape(ctx.world, -1, E, N)
    dt_all = allv[:, :, E * N:]
    h0 = compose_carry(h_all, dt_all, A2.index_select(0, pset_of_job.long()), rev_of_job, ctx.rank)
    return (h0, dt_all.contiguous()) if return_dtsum else h0


# ---- backward (sequence-sharded TRAINING) -------------------------------------------------------------------------
# The adjoint e = dLoss/dh runs against the job's direction, so the roles swap: rank k's adjoint carry-in comes from
# its logical SUCCESSORS.  With Dh_j = the gradient w.r.t. shard j's carry-in from shard j's own tokens
# (csrc/scan_adjoint.cu) and the same per-shard decays exp2(A2 * sum dt_i) as the forward,
#     dhlast_k = sum_{j after k} exp2(A2 * sum_{i between k and j} sum dt_i) * Dh_j
# which is `compose_carry` with the direction flags inverted.  Parameter gradients are partial sums over the local
# tokens: all-reduce them like DDP does (`all_reduce_grads`).
def gather_adjoint(dh_local, dtsum_all, A2, pset_of_job, rev_of_job, ctx):
    """All-gather Dh and compose this rank's adjoint carry-in dhlast (njobs, E, N)."""
    dh_all = _all_gather(dh_local, ctx)                                          # (world, njobs, E, N)
    return compose_carry(dh_all, dtsum_all, A2.index_select(0, pset_of_job.long()), 1 - rev_of_job.to(torch.int32),
                         ctx.rank)


def exchange_halo_grad(dhalo, dx, L, seq_of_job, rev_of_job, ctx):
    """Backward of `gather_halo`: dhalo (njobs, E, 3) is this rank's gradient w.r.t. the 3 samples it borrowed from
    its logical predecessor; add what the logical successor computed for OUR edge samples into dx (njobs, E, >=L)."""
    from . import functional as CF
    host = CF.JOB_HOST.get(seq_of_job.data_ptr())
    rev_l = host[2] if host is not None else tuple(rev_of_job.tolist())
    allv = _all_gather(dhalo, ctx)                                                # (world, njobs, E, 3)
    # a left-to-right job's successor is rank+1 and borrowed our LAST three (same order); a right-to-left job's
    # successor is rank-1 and borrowed our FIRST three in reversed order
    fwd_ok, rev_ok = ctx.rank + 1 < ctx.world, ctx.rank - 1 >= 0
    fwd_jobs = [j for j, r in enumerate(rev_l) if not r]
    rev_jobs = [j for j, r in enumerate(rev_l) if r]
    if fwd_ok and fwd_jobs:
        dx[fwd_jobs, :, L - 3:L] += allv[ctx.rank + 1, fwd_jobs].to(dx.dtype)
    if rev_ok and rev_jobs:
        dx[rev_jobs, :, 0:3] += allv[ctx.rank - 1, rev_jobs].flip(-1).to(dx.dtype)
    return dx


def all_reduce_grads(module, ctx=None, group=None):
    """Sum the parameter gradients over the sequence shards (each rank holds the partial sum over its own tokens)."""
    group = ctx.group if ctx is not None else group
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    if flat.is_cuda and dist.get_backend(group) == "gloo":
        host = flat.cpu()
        dist.all_reduce(host, group=group)
        flat = host.to(flat.device)
    else:
        dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
