"""CPU check of scan variants 9 / 10 (16-bit B/C tile) and 11 / 12 (fp32 tile shared by up to 14 warps) — mbarrier +
arrival-counter hand-over of two tile buffers, no replay pass, optional exp2 software pipeline — through the SIMT
emulation of tests/emu/ — the kernel source caduceus_b200/csrc/scan_fwd_v9.cuh compiled for the host — against the float64 restatement at the kernel boundary
(tests/scan_boundary_ref.py), INCLUDING the sharding / training hooks: conv halo, carry-in state, end state, sum dt,
saved chunk states and the state-only pass.  Not a product path: the product runs only the CUDA build of this source."""
import ctypes as C

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib
from scan_boundary_ref import _problem, boundary_ref
from test_emu_scan_v4 import emu  # noqa: F401  (module-scoped fixture: builds tests/emu/libemu_scan.so)


def _run(lib, L, E, spec, dtype, G, seed, pipe, hooks=False, state_only=False, tile32=False, dt_ready=False):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    njobs, N = len(spec), 16
    delta_f = delta.float()
    if dt_ready:
        # what cad_conv_xproj_fwd writes when dt_b is set: softplus(dt_raw + bias) as FP16 bits, whatever the io dtype
        dt16 = torch.nn.functional.softplus(delta.float() + dt_b[tabs[1].long()][:, :, None]).half()
        delta, delta_f = dt16.view(dtype) if dtype != torch.float16 else dt16, dt16.float()
    g = torch.Generator().manual_seed(seed + 1)
    ld16 = (L + 63) // 64 * 64
    bc16 = torch.zeros(njobs, 2 * N, ld16, dtype=dtype)
    bc16[..., :L] = bc[..., :L].to(dtype)
    bc[..., :L] = bc16[..., :L].float()          # the fp32 tile carries the same (16-bit representable) values
    halo = h0 = None
    if hooks:
        halo = torch.randn(njobs, E, 3, generator=g).to(dtype)
        h0 = torch.randn(njobs, E, N, generator=g)
    nchunks = (L + 511) // 512
    out = torch.full((njobs, E, ld), float("nan")).to(dtype)
    hlast = torch.full((njobs, E, N), float("nan"))
    dtsum = torch.full((njobs, E), float("nan"))
    cstate = torch.full((njobs, E, nchunks, N), float("nan"))
    want = hooks or state_only
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
    a = _lib.ScanFwdArgs(p(xz), p(delta), p(bc) if tile32 else None, None if state_only else p(out),
                         p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), p(halo), p(h0), p(hlast) if want else None,
                         p(dtsum) if want else None, p(cstate) if hooks and not state_only else None,
                         L, E, N, 4, ld, ld, ldbc, ld, xz.shape[0], njobs, conv_w4.shape[0],
                         _lib.CAD_BF16 if dtype == torch.bfloat16 else _lib.CAD_F16, G, int(state_only), 0,
                         (12 if pipe else 11) if tile32 else (10 if pipe else 9), None if tile32 else p(bc16), ld16,
                         int(dt_ready))
    assert lib.emu_scan_v9(C.byref(a), G, int(pipe), int(tile32)) == 0
    f = lambda t: None if t is None else t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), delta_f.numpy(), f(bc16), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk),
                       [s for s, _, _ in spec], [q for _, q, _ in spec], [r for _, _, r in spec], L,
                       halo=f(halo), h0=f(h0), full=True, delta_is_dt=dt_ready)
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    if not state_only:
        got = out.float().numpy()
        assert np.isnan(got[..., L:]).all(), "kernel wrote into the pad columns"
        got = got[..., :L]
        assert np.isfinite(got).all()
        err, bound = np.abs(got - ref[0]), 1e-4 + 1.5 * eps * np.abs(ref[0])
        assert (err <= bound).all(), f"out: max err {err.max():.3e}, worst excess {(err - bound).max():.3e}"
    if want:
        for name, got, r in (("hlast", hlast, ref[1]), ("dtsum", dtsum, ref[2])):
            got = got.numpy()
            assert np.isfinite(got).all(), name
            assert np.allclose(got, r, rtol=2e-4, atol=2e-4 * max(1.0, np.abs(r).max())), (name, np.abs(got - r).max())
    if hooks and not state_only:
        got = cstate.numpy()
        assert np.allclose(got, ref[3], rtol=2e-4, atol=2e-4 * max(1.0, np.abs(ref[3]).max())), np.abs(got - ref[3]).max()
        assert np.array_equal(got[:, :, -1], hlast.numpy()), "last chunk state != end state"


@pytest.mark.parametrize("pipe", [0, 1])
@pytest.mark.parametrize("L", [1, 16, 17, 511, 513, 1030])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v9_ragged_lengths(emu, L, rev, pipe):   # noqa: F811
    _run(emu, L, E=3, spec=[(0, 0, rev)], dtype=torch.bfloat16, G=2, seed=300 + L, pipe=pipe)


@pytest.mark.parametrize("pipe", [0, 1])
def test_emulated_v9_jobs_psets_idle_warps_many_chunks(emu, pipe):   # noqa: F811
    """Caduceus-PS job order, 5 channels over CTAs of 3 warps (one idle warp in the last CTA), five chunks: each tile
    buffer is released and re-armed twice by whichever warp arrives last."""
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    _run(emu, 2300, E=5, spec=spec, dtype=torch.float16, G=3, seed=11, pipe=pipe)


@pytest.mark.parametrize("pipe", [0, 1])
@pytest.mark.parametrize("L", [700, 1537])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v9_hooks(emu, L, rev, pipe):   # noqa: F811
    """conv halo + carry-in state in; end state, sum dt and the per-chunk states out."""
    _run(emu, L, E=4, spec=[(0, 0, rev)], dtype=torch.bfloat16, G=2, seed=500 + L, pipe=pipe, hooks=True)


@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v9_state_only_pass(emu, rev):   # noqa: F811
    _run(emu, 1100, E=4, spec=[(0, 0, rev)], dtype=torch.bfloat16, G=2, seed=77, pipe=1, state_only=True)


@pytest.mark.parametrize("pipe", [0, 1])
@pytest.mark.parametrize("L,rev", [(17, 0), (513, 1), (1030, 0), (1030, 1)])
def test_emulated_v11_fp32_tile_ragged_lengths(emu, L, rev, pipe):   # noqa: F811
    _run(emu, L, E=3, spec=[(0, 0, rev)], dtype=torch.bfloat16, G=2, seed=700 + L, pipe=pipe, tile32=True)


def test_emulated_v12_fp32_tile_wide_cta_with_hooks(emu):   # noqa: F811
    """9 warps in one CTA sharing the fp32 tile (the 14-warp configuration in small), hooks on, four chunks."""
    _run(emu, 1700, E=9, spec=[(0, 0, 0), (0, 0, 1)], dtype=torch.float16, G=9, seed=21, pipe=1, hooks=True, tile32=True)


@pytest.mark.parametrize("tile32", [False, True])
@pytest.mark.parametrize("L,rev", [(513, 0), (1030, 1)])
def test_emulated_v10_v12_dt_precomputed_by_conv_xproj(emu, L, rev, tile32):   # noqa: F811
    """cad_scan_fwd_args.delta_is_dt: `delta` holds dt = softplus(dt_raw + b) as fp16 bits under bf16 I/O (what
    cad_conv_xproj_fwd writes when dt_b is set); the kernel's prologue skips the softplus."""
    _run(emu, L, E=3, spec=[(0, 0, rev), (0, 1, 1 - rev)], dtype=torch.bfloat16, G=2, seed=810 + L, pipe=1, tile32=tile32,
         dt_ready=True)
