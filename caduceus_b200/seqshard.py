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
package) treats its input as the local shard.
"""
import contextlib

import torch
import torch.distributed as dist

_CTX = None


class ShardContext:
    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)


def current():
    return _CTX


@contextlib.contextmanager
def sequence_parallel(group=None):
    """Treat the sequence axis of every BiMamba call inside as sharded over `group` (default: WORLD)."""
    global _CTX
    prev, _CTX = _CTX, ShardContext(group)
    try:
        yield _CTX
    finally:
        _CTX = prev


def _all_gather(t, ctx):
    out = [torch.empty_like(t) for _ in range(ctx.world)]
    dist.all_gather(out, t.contiguous(), group=ctx.group)
    return torch.stack(out)                                    # (world, ...)


def gather_halo(x_rows, L, seq_of_job, rev_of_job, ctx):
    """x_rows (nseq, E, >=L): the conv input rows of this shard.  Returns halo (njobs, E, 3) = the 3 samples
    logically preceding the shard for each job (zeros at the ends of the full sequence), in logical order."""
    assert L >= 3, "sequence shards must hold at least 3 tokens"
    edges = torch.stack([x_rows[..., 0:3], x_rows[..., L - 3:L]])              # (2, nseq, E, 3): first3, last3
    allv = _all_gather(edges, ctx)                                              # (world, 2, nseq, E, 3)
    zero = torch.zeros_like(edges[0, 0])
    halos = []
    from . import functional as CF
    host = CF.JOB_HOST.get(seq_of_job.data_ptr())
    seq_l, rev_l = (host[0], host[2]) if host is not None else (seq_of_job.tolist(), rev_of_job.tolist())
    for s, r in zip(seq_l, rev_l):
        if not r:      # left-to-right: predecessor rank's LAST three samples, already in logical order
            halos.append(allv[ctx.rank - 1, 1, s] if ctx.rank > 0 else zero)
        else:          # right-to-left: successor rank's FIRST three, logical order = physical order reversed
            halos.append(allv[ctx.rank + 1, 0, s].flip(-1) if ctx.rank + 1 < ctx.world else zero)
    return torch.stack(halos).contiguous()


def compose_carry(h_all, dtsum_all, A2_job, rev_of_job, rank):
    """h_all (world, njobs, E, N) zero-carry end states, dtsum_all (world, njobs, E), A2_job (njobs, E, N).
    Carry-in of `rank`:  walk the logical predecessors applying  h <- exp2(A2 * sum dt) * h + H."""
    world = h_all.shape[0]
    fwd = torch.zeros_like(h_all[0])
    for j in range(0, rank):
        fwd = torch.exp2(A2_job * dtsum_all[j][..., None]) * fwd + h_all[j]
    bwd = torch.zeros_like(h_all[0])
    for j in range(world - 1, rank, -1):
        bwd = torch.exp2(A2_job * dtsum_all[j][..., None]) * bwd + h_all[j]
    rev = rev_of_job.to(torch.bool)[:, None, None]
    return torch.where(rev, bwd, fwd).contiguous()


def gather_carry(hlast, dtsum, A2, pset_of_job, rev_of_job, ctx):
    """All-gather (H, sum dt) of the zero-carry pass and compose this rank's carry-in state (njobs, E, N)."""
    E, N = hlast.shape[1], hlast.shape[2]
    packed = torch.cat([hlast.reshape(hlast.shape[0], -1), dtsum], dim=1)       # one message per rank
    allv = _all_gather(packed, ctx)
    h_all = allv[:, :, :E * N].reshape(ctx.world, -1, E, N)
    dt_all = allv[:, :, E * N:]
    return compose_carry(h_all, dt_all, A2.index_select(0, pset_of_job.long()), rev_of_job, ctx.rank)
