"""GPU parity of the lane = channel forward scan (cad_scan_fwd_args.variant = 20: token-major B / C, zero-carry segment scans,
carry composition, segment fix-up) through the C-ABI: against the float64 restatement at the kernel boundary
(tests/scan_boundary_ref.py — the same checker the CPU emulation of this kernel is held to), against the time-parallel kernel
(variant 3) on identical inputs, behind the sharding hooks, and end to end through the model against the fixture produced by the
reference's own code — with both kernels forced in turn."""
import numpy as np
import pytest
import torch

from conftest import golden, tol
from scan_boundary_ref import _problem, boundary_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _run(L, E, spec, dtype, G, seed, variant, **kw):
    from caduceus_b200 import functional as CF
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    bc = bc.to(dtype).float()      # values every variant (fp32 or 16-bit tile) represents exactly
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    out, _, _, _ = CF.scan_fwd(d(xz), d(delta), d(bc), tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)),
                               tuple(d(t) for t in tabs), L, channels_per_cta=G, variant=variant, **kw)
    torch.cuda.synchronize()
    f = lambda t: t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk),
                       [s for s, _, _ in spec], [q for _, q, _ in spec], [r for _, _, r in spec], L)
    return out[..., :L].float().cpu().numpy(), ref


def _check(got, ref, dtype, what):
    # bf16: tighter than the reference's own 3e-2 / 5e-2 (ref:caduceus/tests/test_rcps.py:36); fp16: the reference's.
    # (the CPU emulation of the same source is held to 1.5 output ulps; here MUFU tanh/ex2/lg2 approximations add noise)
    rtol, atol = (1e-2, 1e-2) if dtype == torch.bfloat16 else tol(dtype)
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    assert np.isfinite(got).all(), what
    assert (err <= bound).all(), f"{what}: max err {err.max():.3e}, worst excess {(err - bound).max():.3e}"


@pytest.mark.parametrize("nseg", [1, 3, 7])
@pytest.mark.parametrize("L", [1, 257, 2300, 9000])
@pytest.mark.parametrize("rev", [0, 1])
def test_v20_lane_per_channel_vs_boundary_restatement(L, rev, nseg):
    """variant 20 end to end (token-major B / C copy, zero-carry segment scans, carry composition, segment fix-up) on the
    same checker; nseg = 7 leaves empty blocks at the short lengths."""
    got, ref = _run(L, 96, [(0, 0, rev), (0, 1, 1 - rev)], torch.bfloat16, 0, 600 + L, 20, nseg=nseg)
    # two roundings to bf16 where a carry term is added to an already rounded partial result
    err, bound = np.abs(got - ref), 1e-2 + 1.5e-2 * np.abs(ref)
    assert np.isfinite(got).all() and (err <= bound).all(), (err.max(), (err - bound).max())


@pytest.mark.parametrize("nseg", [1, 4])
@pytest.mark.parametrize("rev", [0, 1])
def test_v20_as_a_sequence_shard(rev, nseg):
    """variant 20 behind the sharding hooks: conv halo in, zero-carry end state + sum dt out (composed from
    the segments), then the shard's carry-in applied to every segment by ONE fix-up launch."""
    from caduceus_b200 import functional as CF
    L, E, dtype = 2300, 96, torch.bfloat16
    spec = [(0, 0, rev), (0, 1, 1 - rev)]
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, 41)
    bc = bc.to(dtype).float()
    g = torch.Generator().manual_seed(2)
    halo, h0 = torch.randn(2, E, 3, generator=g).to(dtype), torch.randn(2, E, 16, generator=g)
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    packed, jobs = tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)), tuple(d(t) for t in tabs)
    out, hl, ds, ctx = CF.scan_fwd(d(xz), d(delta), d(bc), packed, jobs, L, halo=d(halo), want_state=True, variant=20, nseg=nseg)
    f = lambda t: t.float().numpy()   # noqa: E731
    args = (f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk), [0, 0], [0, 1], [rev, 1 - rev], L)
    zero = boundary_ref(*args, halo=f(halo), full=True)
    assert np.allclose(hl.cpu().numpy(), zero[1], rtol=2e-3, atol=2e-3 * max(1.0, np.abs(zero[1]).max()))
    assert np.allclose(ds.cpu().numpy(), zero[2], rtol=2e-3, atol=1e-3)
    CF.scan_fixup(d(xz), d(delta), d(bc), out, packed, jobs, L, d(h0), seg_ctx=ctx)
    ref = boundary_ref(*args, halo=f(halo), h0=f(h0))
    got = out[..., :L].float().cpu().numpy()
    err, bound = np.abs(got - ref), 1e-2 + 1.5e-2 * np.abs(ref)
    assert np.isfinite(got).all() and (err <= bound).all(), (err.max(), (err - bound).max())


def test_v20_helpers_transpose_and_carry_composition():
    from caduceus_b200 import _lib, functional as CF
    import ctypes as C
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    njobs, L, E, nseg = 3, 700, 40, 5
    bc = torch.randn(njobs, 32, 704, generator=g).to(DEV)
    bcT = torch.full((njobs, 768, 32), float("nan"), device=DEV)
    p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    _lib.check(lib.cad_bc_transpose(p(bc), p(bcT), njobs, 32, L, 704, CF._stream()), "cad_bc_transpose")
    assert torch.equal(bcT[:, :L], bc[..., :L].transpose(1, 2)) and (bcT[:, L:] == 0).all()
    st, ds = torch.randn(njobs, nseg, E, 16, generator=g).to(DEV), torch.rand(njobs, nseg, E, generator=g).to(DEV)
    A2 = -(torch.rand(2, E, 16, generator=g) * 8).to(DEV)
    pset = torch.tensor([0, 1, 0], dtype=torch.int32, device=DEV)
    carry = torch.empty(njobs, nseg, E, 16, device=DEV)
    h0 = torch.randn(njobs, E, 16, generator=g).to(DEV)
    hlast, dtsum = torch.empty(njobs, E, 16, device=DEV), torch.empty(njobs, E, device=DEV)
    _lib.check(lib.cad_seg_carry(p(st), p(ds), p(A2), p(pset), p(h0), p(carry), p(hlast), p(dtsum), njobs, nseg, E,
                                 CF._stream()), "cad_seg_carry")
    h = h0.double()
    for s_ in range(nseg):
        assert torch.allclose(carry[:, s_].double(), h, rtol=1e-4, atol=1e-5), s_
        h = torch.exp2(A2[pset.long()].double() * ds[:, s_, :, None].double()) * h + st[:, s_].double()
    assert torch.allclose(hlast.double(), h, rtol=1e-4, atol=1e-5) and torch.allclose(dtsum, ds.sum(1), rtol=1e-5)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_v20_agrees_with_v3_on_identical_inputs(dtype):
    """Caduceus-PS job table, 12 segments of 2048+ tokens: the two kernels on the same device buffers."""
    from caduceus_b200 import functional as CF
    L, E = 30000, 128
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, 77)
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    args = (d(xz), d(delta), d(bc), tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)), tuple(d(t) for t in tabs), L)
    o3 = CF.scan_fwd(*args, variant=3)[0][..., :L].float()
    o20 = CF.scan_fwd(*args, variant=20, nseg=12)[0][..., :L].float()
    rtol, atol = tol(dtype)
    assert torch.isfinite(o20).all()
    assert torch.all((o20 - o3).abs() <= atol * 0.2 + rtol * 0.5 * o3.abs()), (o20 - o3).abs().max()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_xproj_token_major_copy(dtype):
    """cad_conv_xproj_args.bcT: the B / C rows token-major (what variant 20 reads) == the transpose of the state-major rows the
    same launch writes; zeros beyond L."""
    from caduceus_b200 import functional as CF
    L, E, R, N = 1500, 128, 8, 16
    g = torch.Generator().manual_seed(5)
    xz = torch.randn(2, 2 * E, 1504, generator=g).to(dtype).to(DEV)
    w_x = (torch.randn(2, R + 2 * N, E, generator=g) * E ** -0.5).to(dtype).to(DEV)
    w_dt = (torch.randn(2, E, R, generator=g) * R ** -0.5).to(dtype).to(DEV)
    conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(DEV)
    conv_b = (0.1 * torch.randn(2, E, generator=g)).to(DEV)
    jobs = tuple(torch.tensor(v, dtype=torch.int32, device=DEV) for v in ([0, 0, 1, 1], [0, 1, 0, 1], [0, 1, 1, 0]))
    raw, bc0 = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L)
    raw2, bc2, bcT = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=True)
    assert torch.equal(bc0, bc2) and torch.equal(raw[..., :L], raw2[..., :L]) and bcT.shape == (4, 1536, 2 * N)
    assert torch.equal(bcT[:, :L], bc0[..., :L].transpose(1, 2)) and (bcT[:, L:] == 0).all()


@pytest.mark.parametrize("variant", [3, 20])
@pytest.mark.parametrize("tag", ["ps_small", "ph_config0"])
def test_model_forward_with_scan_variant_vs_reference_fixture(tag, variant):
    """The whole model with the forward scan forced to either kernel, against the logits the reference's own code produced."""
    import caduceus
    from caduceus_b200 import functional as CF
    fx = golden(f"model_{tag}.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = caduceus.CaduceusForMaskedLM(cfg)
    model.load_state_dict(fx["state_dict"])
    model = model.to(DEV).to(torch.bfloat16).eval()
    launches = []
    orig = CF.choose_scan_variant
    try:
        CF.SCAN_VARIANT = variant
        CF.choose_scan_variant = lambda *a: launches.append(orig(*a)) or launches[-1]
        with torch.no_grad():
            logits = model(fx["input_ids"].to(DEV)).logits.float().cpu()
    finally:
        CF.SCAN_VARIANT = 0
        CF.choose_scan_variant = orig
    assert launches and all(v == variant for v in launches), launches
    # same criterion as test_gpu_parity.py::test_model_low_precision_vs_reference_fixture (16-bit stack vs fp32 fixture)
    rtol, atol = tol(torch.bfloat16)
    err = (logits - fx["logits"]).abs()
    assert torch.all(err <= atol + rtol * fx["logits"].abs().max()), err.max()
