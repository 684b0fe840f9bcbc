"""CPU check of the carry fix-up kernel (caduceus_b200/csrc/scan_fixup.cuh) through the SIMT emulation of tests/emu/ — the
kernel source compiled for the host:
  * whole-sequence mode (the multi-GPU path): zero-carry output + fix-up(h0) == the operator with carry-in h0;
  * segment mode (the last stage of scan variant 20): the emulated lane = channel kernel's zero-carry segment outputs, the
    carries composed as cad_seg_carry does, and the emulated fix-up reproduce the UNSEGMENTED operator — for reversed jobs,
    segment lengths that are not multiples of the fix-up's 512-token chunks, ragged tails and empty blocks.
Not a product path: the product runs only the CUDA build of these sources."""
import ctypes as C

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib
from scan_boundary_ref import _problem, boundary_ref
from test_emu_scan_v4 import emu  # noqa: F401  (module-scoped fixture: builds tests/emu/libemu_scan.so)

N = 16
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
f = lambda t: t.float().numpy()   # noqa: E731


def _io(dtype):
    return {torch.bfloat16: _lib.CAD_BF16, torch.float16: _lib.CAD_F16, torch.float32: _lib.CAD_F32}[dtype]


def _fixup(lib, prob, out, L, E, njobs, dtype, G, h0=None, nseg=0, carry=None, cutoff=-40.0, seg_first=0):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = prob
    a = _lib.ScanFixupArgs(p(xz), p(delta), p(bc), p(out), p(dt_b), p(A2), p(tabs[0]), p(tabs[1]), p(tabs[2]), p(h0),
                           L, E, N, ld, ld, ldbc, ld, xz.shape[0], njobs, _io(dtype), G, cutoff, nseg, p(carry), seg_first)
    assert lib.emu_scan_fixup(C.byref(a), G) == 0


def _bound(ref, partial, dtype):
    eps = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 2.0 ** -22}[dtype]
    # two roundings (the zero-carry partial result, then the sum): up to one ulp, and an ulp is up to 2 eps |x|
    return 2e-4 + 2.1 * eps * np.maximum(np.abs(ref), np.abs(partial))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("L", [17, 700, 1537])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_fixup_whole_sequence_mode(emu, L, rev, dtype):   # noqa: F811
    E, spec = 5, [(0, 0, rev), (0, 1, 1 - rev)]
    prob = _problem(L, E, spec, dtype, 4000 + L)
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = prob
    g = torch.Generator().manual_seed(L)
    h0 = torch.randn(2, E, N, generator=g)
    h0[0, 1] = 0.0                                                    # a channel without carry: nothing to add
    args = (f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk), [0, 0], [0, 1], [rev, 1 - rev], L)
    zero = boundary_ref(*args)
    ref = boundary_ref(*args, h0=f(h0))
    out = torch.full((2, E, ld), float("nan")).to(dtype)
    out[..., :L] = torch.from_numpy(zero).to(dtype)
    part = out[..., :L].float().numpy().copy()
    _fixup(emu, prob, out, L, E, 2, dtype, G=3, h0=h0)
    got = out.float().numpy()
    assert np.isnan(got[..., L:]).all(), "kernel wrote into the pad columns"
    err = np.abs(got[..., :L] - ref)
    assert (err <= _bound(ref, part, dtype)).all(), err.max()


@pytest.mark.parametrize("cutoff", [-24.0, -40.0])
@pytest.mark.parametrize("L,nseg", [(1100, 2), (1537, 3), (1100, 5), (3000, 4)])
def test_emulated_v20_pipeline_pass_a_carry_fixup(emu, L, nseg, cutoff):   # noqa: F811
    """variant 20 end to end on CPU: emulated segment scans -> carries -> emulated segment-mode fix-up == the operator."""
    _pipeline(emu, L, nseg, cutoff, shard=False)


@pytest.mark.parametrize("L,nseg", [(5, 2), (700, 2), (1537, 3), (1100, 5), (767, 3), (1022, 2)])   # 767, 1022: fewer than 3 masked tail tokens
def test_emulated_v20_pipeline_as_a_sequence_shard(emu, L, nseg):   # noqa: F811
    """The same pipeline as ONE SHARD of a longer sequence: conv halo into the segment scans, the shard's
    carry-in h0 into the carry composition, every segment (the first included) fixed up in one launch; the composed end
    state and sum dt are what the shard hands to its neighbours."""
    _pipeline(emu, L, nseg, -24.0, shard=True)


def _pipeline(emu, L, nseg, cutoff, shard, E=40, W=2, variant=20):   # noqa: F811
    dtype = torch.bfloat16
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    prob = _problem(L, E, spec, dtype, 70 + L + nseg)
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = prob
    njobs = len(spec)
    halo = h0 = None
    if shard:
        g = torch.Generator().manual_seed(L + nseg)
        halo = torch.randn(njobs, E, 3, generator=g).to(dtype)
        h0 = torch.randn(njobs, E, N, generator=g)
    Lp = (L + 255) // 256 * 256
    bcT = torch.zeros(njobs, Lp, 2 * N)
    bcT[:, :L] = bc[..., :L].transpose(1, 2)
    out = torch.full((njobs, E, ld), float("nan")).to(dtype)
    seg_state = torch.full((njobs, nseg, E, N), float("nan"))
    seg_dtsum = torch.full((njobs, nseg, E), float("nan"))
    a = _lib.ScanFwdArgs(p(xz), p(delta), None, p(out), p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), p(halo), None, None, None, None,
                         L, E, N, 4, ld, ld, ldbc, ld, xz.shape[0], njobs, conv_w4.shape[0], _io(dtype), W, 0, 0, variant, None, 0, 0,
                         p(bcT), nseg, p(seg_state), p(seg_dtsum))
    assert emu.emu_scan_v20(C.byref(a), W) == 0
    part = out[..., :L].float().numpy().copy()
    # cad_seg_carry in fp32 (csrc/scan_fwd_v20.cu::seg_carry_kernel)
    carry = torch.empty(njobs, nseg, E, N)
    h = torch.zeros(njobs, E, N) if h0 is None else h0.clone()
    a2 = A2[tabs[1].long()]
    for s in range(nseg):
        carry[:, s] = h
        h = torch.exp2(a2 * seg_dtsum[:, s, :, None]) * h + seg_state[:, s]
    _fixup(emu, prob, out, L, E, njobs, dtype, G=7, nseg=nseg, carry=carry.contiguous(), cutoff=cutoff, seg_first=int(shard))
    ref = boundary_ref(f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk),
                       [s for s, _, _ in spec], [q for _, q, _ in spec], [r for _, _, r in spec], L,
                       halo=None if halo is None else f(halo), h0=None if h0 is None else f(h0), full=True)
    if shard:      # what the shard reports: end state with its own carry-in folded in, and the total sum dt
        assert np.allclose(h.numpy(), ref[1], rtol=2e-4, atol=2e-4 * max(1.0, np.abs(ref[1]).max()))
        assert np.allclose(seg_dtsum.sum(1).numpy(), ref[2], rtol=2e-4, atol=1e-4)
    ref = ref[0]
    got = out.float().numpy()
    assert np.isnan(got[..., L:]).all(), "a kernel wrote into the pad columns"
    err = np.abs(got[..., :L] - ref)
    # dropped terms: below 2^cutoff of |C h0| — far under the bound for both cut-offs
    assert (err <= _bound(ref, part, dtype)).all(), (err.max(), (err - _bound(ref, part, dtype)).max())
