"""CPU check of the second backward-scan kernel (caduceus_b200/csrc/scan_bwd_v2.cuh: 8 tokens per lane, two 256-token
passes per saved chunk, ehat-form adjoint, 128-register budget) through the SIMT emulation of tests/emu/ — the kernel
source compiled for the host — against float64 autograd of the operator at the kernel boundary
(tests/scan_boundary_ref.py::boundary_grads).  Covers ragged lengths, reversed jobs, the Caduceus-PS job table with idle
warps (cross-warp dB / dC sums, parameter-set atomics), and the sharding hooks (conv halo, carry-in state h0 -> dh0,
adjoint carry-in dhlast).  Not a product path: the product runs only the CUDA build of this source."""
import ctypes as C

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib
from scan_boundary_ref import _problem, boundary_grads
from test_emu_scan_v4 import emu  # noqa: F401  (module-scoped fixture: builds tests/emu/libemu_scan.so)


def _run(lib, L, E, spec, dtype, G, seed, hooks=False):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    njobs, N, P = len(spec), 16, conv_w4.shape[0]
    g = torch.Generator().manual_seed(seed + 1)
    dout = torch.randn(njobs, E, ld, generator=g).to(dtype)
    dout[..., L:] = 5.0                                          # junk in the pad columns must not leak
    halo = h0 = dhlast = None
    if hooks:
        halo = torch.randn(njobs, E, 3, generator=g).to(dtype)
        h0 = torch.randn(njobs, E, N, generator=g)
        dhlast = torch.randn(njobs, E, N, generator=g)
    seq, pset, rev = ([s[k] for s in spec] for k in range(3))
    ref = boundary_grads(xz.float(), delta.float(), bc, dout.float(), conv_w4, conv_b, dt_b, A2, Dk, seq, pset, rev, L,
                         halo=None if halo is None else halo.float(), h0=h0, dhlast=dhlast)
    cstate = ref["chunk_state"].float().contiguous()             # what the forward kernel saved
    nan = float("nan")
    dz = torch.full((njobs, E, ld), nan).to(dtype)
    du = torch.full((njobs, E, ld), nan).to(dtype)
    dd = torch.full((njobs, E, ld), nan).to(dtype)
    dbc = torch.zeros(njobs, 2 * N, ldbc)
    ddt_b, dA2, dD = torch.zeros(P, E), torch.zeros(P, E, N), torch.zeros(P, E)
    dh0 = torch.full((njobs, E, N), nan) if hooks else None
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
    io = {torch.bfloat16: _lib.CAD_BF16, torch.float16: _lib.CAD_F16, torch.float32: _lib.CAD_F32}[dtype]
    a = _lib.ScanBwdArgs(p(xz), p(delta), p(bc), p(dout), p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), p(halo), p(h0), p(cstate),
                         p(dz), p(du), p(dd), p(dbc), p(ddt_b), p(dA2), p(dD), p(dh0), p(dhlast),
                         L, E, N, 4, ld, ld, ldbc, ld, ld, ld, ld, xz.shape[0], njobs, P, io, G)
    assert lib.emu_scan_bwd_v2(C.byref(a), G) == 0

    eps = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 2.0 ** -22}[dtype]

    def close(name, got, want, out_eps, rel=2e-4):
        got, want = got.double().numpy(), want.double().numpy()
        assert np.isfinite(got).all(), name
        scale = max(1.0, np.abs(want).max())
        err, bound = np.abs(got - want), rel * scale + (out_eps * 1.5 + rel) * np.abs(want)
        assert (err <= bound).all(), f"{name}: max err {err.max():.3e} (scale {scale:.3e}), worst excess {(err - bound).max():.3e}"

    for name, got, want in (("dz", dz, ref["dz"]), ("du", du, ref["du"]), ("ddelta", dd, ref["ddelta"])):
        g_ = got.float()
        assert torch.isnan(g_[..., L:]).all(), f"{name}: kernel wrote into the pad columns"
        close(name, g_[..., :L], want, eps)
    assert (dbc[..., L:] == 0).all(), "dbc: pad columns must stay zero"
    close("dbc", dbc[..., :L], ref["dbc"], 0.0)
    close("ddt_b", ddt_b, ref["ddt_b"], 0.0)
    close("dA2", dA2, ref["dA2"], 0.0)
    close("dD", dD, ref["dD"], 0.0)
    if hooks:
        close("dh0", dh0, ref["dh0"], 0.0)


@pytest.mark.parametrize("L", [1, 8, 9, 255, 257, 513, 1030])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_bwd_v2_ragged_lengths(emu, L, rev):   # noqa: F811
    _run(emu, L, E=3, spec=[(0, 0, rev)], dtype=torch.float32, G=2, seed=900 + L)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_emulated_bwd_v2_jobs_psets_idle_warps(emu, dtype):   # noqa: F811
    """Caduceus-PS job order, 5 channels over CTAs of 3 warps (one idle warp in the last CTA), three chunks."""
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    _run(emu, 1300, E=5, spec=spec, dtype=dtype, G=3, seed=13)


@pytest.mark.parametrize("L", [700, 1537])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_bwd_v2_hooks(emu, L, rev):   # noqa: F811
    """conv halo + carry-in state + adjoint carry-in (dhlast) in; dh0 out."""
    _run(emu, L, E=4, spec=[(0, 0, rev)], dtype=torch.float32, G=2, seed=950 + L, hooks=True)
