"""CPU check of scan variant 20 (lane = channel: a warp owns 32 channels and walks time serially; the sequence is cut
into `nseg` segments per job, each scanned from a zero state) through the SIMT emulation of tests/emu/ — the kernel source
caduceus_b200/csrc/scan_fwd_v20.cuh compiled for the host — against the float64 restatement at the kernel boundary
(tests/scan_boundary_ref.py).  With nseg > 1 the kernel's three outputs (zero-carry outputs, end state and sum dt of every
segment) are pushed through a float64 restatement of what cad_seg_carry + cad_bimamba_scan_fixup do, and the result must
equal the unsegmented operator: that pins the segment geometry (physical blocks, logical order of reversed jobs, ragged
tails, empty blocks) and the carry algebra.  Not a product path: the product runs only the CUDA build of this source."""
import ctypes as C

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib
from scan_boundary_ref import _problem, _silu, _softplus, boundary_ref
from emu_build import emu  # noqa: F401  (module-scoped fixture: builds tests/emu/libemu_scan.so)

N, CH = 16, 256


def _blocks(L, nseg):
    nch = (L + CH - 1) // CH
    per = (nch + nseg - 1) // nseg
    return [(min(k * per * CH, L), min((k + 1) * per * CH, L)) for k in range(nseg)]


def _apply_carries(out0, seg_state, seg_dtsum, xz, delta, bc, dt_b, A2, spec, L, nseg, delta_is_dt):
    """float64 restatement of cad_seg_carry + cad_bimamba_scan_fixup on the kernel's outputs."""
    out = out0.astype(np.float64).copy()
    E = delta.shape[1]
    for j, (s, p, rev) in enumerate(spec):
        a2 = A2[p].astype(np.float64)
        h = np.zeros((E, N))
        for sl in range(nseg):
            lo, hi = _blocks(L, nseg)[nseg - 1 - sl if rev else sl]
            if hi > lo and sl > 0:
                idx = np.arange(hi - 1, lo - 1, -1) if rev else np.arange(lo, hi)       # logical order
                dr = delta[j][:, idx].astype(np.float64)
                dt = dr if delta_is_dt else _softplus(dr + dt_b[p].astype(np.float64)[:, None])
                cum = np.cumsum(dt, axis=1)                                             # (E, T)
                Cm = bc[j, N:][:, idx].astype(np.float64)                               # (N, T)
                z = xz[s, E:][:, idx].astype(np.float64)
                term = np.einsum("nt,ent,en->et", Cm, np.exp2(a2[:, :, None] * cum[:, None, :]), h)
                out[j][:, idx] += term * _silu(z)
            h = np.exp2(a2 * seg_dtsum[j, sl][:, None]) * h + seg_state[j, sl]
    return out


def _run(lib, L, E, spec, dtype, W, seed, nseg, small_dt=False):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    if small_dt:            # every dt below 0.118: the warp-uniform series form of the softplus (no lg2) is taken
        delta = (delta.float() * 0.05).to(dtype)
        dt_b[:, 0] = dt_b[:, 1]
    njobs = len(spec)
    delta_f = delta.float()
    Lp = (L + CH - 1) // CH * CH
    bcT = torch.zeros(njobs, Lp, 2 * N)
    bcT[:, :L] = bc[..., :L].transpose(1, 2)
    out = torch.full((njobs, E, ld), float("nan")).to(dtype)
    seg_state = torch.full((njobs, nseg, E, N), float("nan"))
    seg_dtsum = torch.full((njobs, nseg, E), float("nan"))
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
    a = _lib.ScanFwdArgs(p(xz), p(delta), None, p(out), p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), None, None, None, None, None,
                         L, E, N, 4, ld, ld, ldbc, ld, xz.shape[0], njobs, conv_w4.shape[0],
                         _lib.CAD_BF16 if dtype == torch.bfloat16 else _lib.CAD_F16, W, 0, 20,
                         p(bcT), nseg, p(seg_state), p(seg_dtsum))
    assert lib.emu_scan_v20(C.byref(a), W) == 0
    f = lambda t: t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), delta_f.numpy(), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk),
                       [s for s, _, _ in spec], [q for _, q, _ in spec], [r for _, _, r in spec], L, full=True)
    got0 = out.float().numpy()
    assert np.isnan(got0[..., L:]).all(), "kernel wrote into the pad columns"
    got0 = got0[..., :L]
    assert np.isfinite(got0).all() and np.isfinite(seg_state.numpy()).all() and np.isfinite(seg_dtsum.numpy()).all()
    got = _apply_carries(got0, seg_state.numpy().astype(np.float64), seg_dtsum.numpy().astype(np.float64), f(xz),
                         delta_f.numpy(), f(bc), f(dt_b), f(A2), spec, L, nseg, False)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    # the zero-carry output was rounded to the I/O dtype BEFORE the carry term is added (as on the device): 1.5 ulps of
    # the partial result; with one segment there is no carry term and the bound is the usual one
    scale = np.abs(ref[0]) if nseg == 1 else np.maximum(np.abs(ref[0]), np.abs(got0))
    err, bound = np.abs(got - ref[0]), 1e-4 + 1.5 * eps * scale
    assert (err <= bound).all(), f"out: max err {err.max():.3e}, worst excess {(err - bound).max():.3e}"
    # the composed end state and the total sum dt equal the unsegmented ones
    for j, (s, q, rev) in enumerate(spec):
        h = np.zeros((E, N))
        for sl in range(nseg):
            h = np.exp2(f(A2)[q].astype(np.float64) * seg_dtsum.numpy()[j, sl][:, None]) * h + seg_state.numpy()[j, sl]
        assert np.allclose(h, ref[1][j], rtol=2e-4, atol=2e-4 * max(1.0, np.abs(ref[1][j]).max()))
        assert np.allclose(seg_dtsum.numpy()[j].sum(0), ref[2][j], rtol=2e-4, atol=1e-4)


@pytest.mark.parametrize("L", [1, 7, 8, 9, 255, 257, 700])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v20_one_segment_ragged_lengths(emu, L, rev):   # noqa: F811
    _run(emu, L, E=40, spec=[(0, 0, rev)], dtype=torch.bfloat16, W=2, seed=2000 + L, nseg=1)


@pytest.mark.parametrize("nseg", [2, 3, 5])
@pytest.mark.parametrize("L", [1100, 1537])
def test_emulated_v20_segments_compose_to_the_unsegmented_scan(emu, L, nseg):   # noqa: F811
    """Caduceus-PS job table (both directions, two parameter sets), 2 warps per CTA with idle lanes (E = 40), several
    chunks per segment, a ragged last block and — for nseg = 5 at L = 1100 — an empty one."""
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    _run(emu, L, E=40, spec=spec, dtype=torch.float16, W=2, seed=31 + nseg, nseg=nseg)


@pytest.mark.parametrize("L", [6, 7, 767, 1022])
def test_emulated_v20_halo_with_fewer_than_three_masked_tail_tokens(emu, L):   # noqa: F811
    """Reversed job + conv halo + a ragged end with 1 or 2 masked tokens in the last 8-token group: the masked tokens carry
    halo values themselves, so the initial conv window must start beyond them (regression: they were entered twice)."""
    import ctypes as C2
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, 8, [(0, 0, 1), (0, 1, 0)], torch.bfloat16, 3100 + L)
    g = torch.Generator().manual_seed(L)
    halo = torch.randn(2, 8, 3, generator=g).to(torch.bfloat16)
    Lp = (L + CH - 1) // CH * CH
    bcT = torch.zeros(2, Lp, 2 * N)
    bcT[:, :L] = bc[..., :L].transpose(1, 2)
    out = torch.full((2, 8, ld), float("nan")).to(torch.bfloat16)
    p = lambda t: None if t is None else C2.c_void_p(t.data_ptr())   # noqa: E731
    a = _lib.ScanFwdArgs(p(xz), p(delta), None, p(out), p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), p(halo), None, None, None, None,
                         L, 8, N, 4, ld, ld, ldbc, ld, xz.shape[0], 2, conv_w4.shape[0], _lib.CAD_BF16, 1, 0, 20,
                         p(bcT), 1, None, None)
    assert emu.emu_scan_v20(C2.byref(a), 1) == 0
    f = lambda t: t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk), [0, 0], [0, 1], [1, 0], L, halo=f(halo))
    got = out.float().numpy()[..., :L]
    err, bound = np.abs(got - ref), 1e-4 + 1.5 * 2.0 ** -8 * np.abs(ref)
    assert (err <= bound).all(), (err.max(), (err - bound).max())


def test_emulated_v20_many_chunks(emu):   # noqa: F811
    _run(emu, 2300, E=64, spec=[(0, 0, 0), (0, 1, 1)], dtype=torch.bfloat16, W=1, seed=5, nseg=2)


@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v20_softplus_series_path(emu, rev):   # noqa: F811
    """init-time dt range in every lane of every warp: dt comes from the log1p series instead of lg2."""
    _run(emu, 1100, E=64, spec=[(0, 0, rev), (0, 1, 1 - rev)], dtype=torch.bfloat16, W=2, seed=91, nseg=2, small_dt=True)
