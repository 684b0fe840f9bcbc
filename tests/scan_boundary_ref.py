"""float64 restatement of the fused scan operator AT THE KERNEL BOUNDARY (test infrastructure).

Inputs are exactly what cad_bimamba_scan_fwd takes (include/caduceus_b200.h, "Semantics in LOGICAL time tau"): the
channel-major in-proj output xz, dt_raw, the B / C rows and the packed parameters; the arithmetic follows upstream
selective_scan_ref + causal_conv1d in float64, token by token.  Used by the CPU
emulation tests (tests/test_emu_scan_*.py) and by the GPU parity tests of the scan variants."""
import numpy as np
import torch


def _softplus(v):
    return np.where(v > 20.0, v, np.log1p(np.exp(np.minimum(v, 20.0))))


def _silu(v):
    return v / (1.0 + np.exp(-v))


def boundary_ref(xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, seq, pset, rev, L, halo=None, h0=None, full=False,
                 chunk=512, delta_is_dt=False):
    """float64 restatement at the kernel boundary.  xz (nseq, 2E, ld), delta (njobs, E, ld), bc (njobs, 2N, ldbc).
    halo (njobs, E, 3): x at logical times -3, -2, -1; h0 (njobs, E, N): carry-in state.
    full=True also returns (hlast, dtsum, chunk_state) — the state after every `chunk` LOGICAL tokens counted the way
    the kernels do (chunks aligned in physical time: a reversed job's first logical chunk is the ragged one)."""
    njobs, E = delta.shape[0], delta.shape[1]
    N = bc.shape[1] // 2
    out = np.zeros((njobs, E, L))
    nchunks = (L + chunk - 1) // chunk
    hlast = np.zeros((njobs, E, N))
    dtsum = np.zeros((njobs, E))
    cstate = np.zeros((njobs, E, nchunks, N))
    for j in range(njobs):
        s, p = seq[j], pset[j]
        idx = np.arange(L)[::-1] if rev[j] else np.arange(L)      # logical time -> physical token
        x = xz[s, :E, :L].astype(np.float64)[:, idx]
        z = xz[s, E:, :L].astype(np.float64)[:, idx]
        dr = delta[j, :, :L].astype(np.float64)[:, idx]
        B = bc[j, :N, :L].astype(np.float64)[:, idx]
        Cm = bc[j, N:, :L].astype(np.float64)[:, idx]
        pre = np.zeros((E, 3)) if halo is None else halo[j].astype(np.float64)
        xp = np.concatenate([pre, x], axis=1)
        w = conv_w4[p].astype(np.float64)
        u = _silu(conv_b[p].astype(np.float64)[:, None] + sum(w[:, k:k + 1] * xp[:, k:k + L] for k in range(4)))
        dt = dr if delta_is_dt else _softplus(dr + dt_b[p].astype(np.float64)[:, None])   # cad_scan_fwd_args.delta_is_dt
        a2 = A2[p].astype(np.float64)                              # (E, N), already * log2(e)
        h = np.zeros((E, N)) if h0 is None else h0[j].astype(np.float64).copy()
        y = np.zeros((E, L))
        first = L - (nchunks - 1) * chunk if rev[j] else chunk      # logical length of the first visited chunk
        bounds = [min(L, first + c * chunk) for c in range(nchunks)]
        ci = 0
        for t in range(L):
            h = np.exp2(dt[:, t:t + 1] * a2) * h + (dt[:, t] * u[:, t])[:, None] * B[None, :, t]
            y[:, t] = (h * Cm[None, :, t]).sum(1) + Dk[p].astype(np.float64) * u[:, t]
            while ci < nchunks and t + 1 == bounds[ci]:
                cstate[j, :, ci] = h
                ci += 1
        o = y * _silu(z)
        out[j][:, idx] = o
        hlast[j] = h
        dtsum[j] = dt.sum(1)
    return (out, hlast, dtsum, cstate) if full else out


def _problem(L, E, njobs_spec, dtype, seed):
    """njobs_spec: list of (seq, pset, rev)."""
    g = torch.Generator().manual_seed(seed)
    N = 16
    nseq = max(s for s, _, _ in njobs_spec) + 1
    npset = max(p for _, p, _ in njobs_spec) + 1
    njobs = len(njobs_spec)
    ld = (L + 15) // 16 * 16
    ldbc = (L + 31) // 32 * 32
    xz = torch.randn(nseq, 2 * E, ld, generator=g).to(dtype)
    xz[..., L:] = 7.0                                           # junk in the pad columns must not leak into [0, L)
    delta = (torch.randn(njobs, E, ld, generator=g) * 1.5).to(dtype)
    delta[..., L:] = 9.0
    bc = torch.zeros(njobs, 2 * N, ldbc)
    bc[..., :L] = torch.randn(njobs, 2 * N, L, generator=g)
    conv_w4 = (0.5 * torch.randn(npset, E, 4, generator=g)).contiguous()
    conv_b = 0.1 * torch.randn(npset, E, generator=g)
    dt_b = torch.log(torch.expm1(torch.exp(torch.rand(npset, E, generator=g) * 4.6 - 6.9)))   # dt in [1e-3, 0.1]
    dt_b[:, 0] = 25.0                                           # exercises the softplus threshold branch
    A2 = (-torch.arange(1, N + 1, dtype=torch.float32).repeat(npset, E, 1)
          * (0.5 + torch.rand(npset, E, 1, generator=g)) * 1.4426950408889634).contiguous()
    Dk = torch.randn(npset, E, generator=g)
    tabs = [torch.tensor([j[k] for j in njobs_spec], dtype=torch.int32) for k in range(3)]
    return xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc




def boundary_grads(xz, delta, bc, dout, conv_w4, conv_b, dt_b, A2, Dk, seq, pset, rev, L, halo=None, h0=None, dhlast=None,
                   chunk=512):
    """float64 autograd of the operator at the kernel boundary = what cad_bimamba_scan_bwd must return
    (include/caduceus_b200.h): dz, du (w.r.t. u = silu(conv(x)) as an independent input: the conv is differentiated by
    a separate kernel), ddelta (w.r.t. dt_raw), dbc, ddt_b, dA2, dD, dh0 — for the loss  sum(out * dout) [+ sum(hlast *
    dhlast)].  Also returns the forward's chunk states (the saved tensor the backward consumes).  torch tensors in."""
    T = torch.float64
    njobs, E = delta.shape[0], delta.shape[1]
    N = bc.shape[1] // 2
    P = A2.shape[0]
    nchunks = (L + chunk - 1) // chunk
    dt_b_l = dt_b.to(T).clone().requires_grad_(True)
    A2_l = A2.to(T).clone().requires_grad_(True)
    D_l = Dk.to(T).clone().requires_grad_(True)
    z_l = [xz[seq[j], E:, :L].to(T).clone().requires_grad_(True) for j in range(njobs)]
    dr_l = [delta[j, :, :L].to(T).clone().requires_grad_(True) for j in range(njobs)]
    bc_l = [bc[j, :, :L].to(T).clone().requires_grad_(True) for j in range(njobs)]
    h0_l = [(torch.zeros(E, N, dtype=T) if h0 is None else h0[j].to(T).clone()).requires_grad_(True) for j in range(njobs)]
    u_l, loss, cstates = [], 0.0, torch.zeros(njobs, E, nchunks, N, dtype=T)
    for j in range(njobs):
        p = pset[j]
        idx = torch.arange(L - 1, -1, -1) if rev[j] else torch.arange(L)       # logical time -> physical token
        x = xz[seq[j], :E, :L].to(T)[:, idx]
        pre = torch.zeros(E, 3, dtype=T) if halo is None else halo[j].to(T)
        xp = torch.cat([pre, x], dim=1)
        w = conv_w4[p].to(T)
        c = conv_b[p].to(T)[:, None] + sum(w[:, k:k + 1] * xp[:, k:k + L] for k in range(4))
        u = (c * torch.sigmoid(c)).detach().clone().requires_grad_(True)           # logical order
        u_l.append(u)
        raw = dr_l[j][:, idx] + dt_b_l[p][:, None]
        dt = torch.where(raw > 20.0, raw, torch.log1p(torch.exp(torch.clamp(raw, max=20.0))))
        B, Cm = bc_l[j][:N][:, idx], bc_l[j][N:][:, idx]
        zz = z_l[j][:, idx]
        h = h0_l[j]
        first = L - (nchunks - 1) * chunk if rev[j] else chunk
        bounds = [min(L, first + cc * chunk) for cc in range(nchunks)]
        ci, ys = 0, []
        for t in range(L):
            h = torch.exp2(dt[:, t:t + 1] * A2_l[p]) * h + (dt[:, t] * u[:, t])[:, None] * B[None, :, t]
            ys.append((h * Cm[None, :, t]).sum(1) + D_l[p] * u[:, t])
            while ci < nchunks and t + 1 == bounds[ci]:
                cstates[j, :, ci] = h.detach()
                ci += 1
        y = torch.stack(ys, dim=1)
        out = y * (zz * torch.sigmoid(zz))
        loss = loss + (out * dout[j, :, :L].to(T)[:, idx]).sum()
        if dhlast is not None:
            loss = loss + (h * dhlast[j].to(T)).sum()
    loss.backward()
    zero = lambda t: torch.zeros_like(t) if t.grad is None else t.grad     # noqa: E731
    inv = lambda j: (torch.arange(L - 1, -1, -1) if rev[j] else torch.arange(L))   # noqa: E731  (self-inverse map)
    dz = torch.stack([zero(z_l[j]) for j in range(njobs)])
    ddelta = torch.stack([zero(dr_l[j]) for j in range(njobs)])
    dbc = torch.stack([zero(bc_l[j]) for j in range(njobs)])
    du = torch.stack([zero(u_l[j])[:, inv(j)] for j in range(njobs)])              # back to physical order
    dh0 = torch.stack([zero(h0_l[j]) for j in range(njobs)])
    return dict(dz=dz, du=du, ddelta=ddelta, dbc=dbc, ddt_b=zero(dt_b_l), dA2=zero(A2_l), dD=zero(D_l), dh0=dh0,
                chunk_state=cstates)
