"""GPU parity of the non-default scan kernels (cad_scan_fwd_args.variant = 4 paired channels, 7 no replay,
9 / 10 16-bit tile + barrier-free hand-over, 11 / 12 the same on a shared fp32 tile; cad_scan_bwd_args.variant = 2, the
two-CTAs-per-SM backward), through the C-ABI: against the float64 restatement at the kernel boundary
(tests/scan_boundary_ref.py — the same checker the CPU emulation of these kernels is held to), against the default
kernel (variant 3) on identical inputs, and end to end through the model against the fixture produced by the
reference's own code."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, tol
from scan_boundary_ref import _problem, boundary_grads, boundary_ref

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


@pytest.mark.parametrize("L", [1, 17, 511, 512, 513, 1030, 2300])
@pytest.mark.parametrize("rev", [0, 1])
def test_v4_vs_boundary_restatement_ragged_lengths(L, rev):
    got, ref = _run(L, 64, [(0, 0, rev)], torch.bfloat16, 0, 100 + L, 4)
    _check(got, ref, torch.bfloat16, f"v4 L={L} rev={rev}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("G", [1, 2, 5, 7])
def test_v4_ps_job_layout_and_cta_shapes(dtype, G):
    """Caduceus-PS job order (2 sequences x 2 parameter sets, rev = direction XOR strand), E/2 = 19 pairs so the last
    CTA has idle warps for every G, five chunks with a ragged tail."""
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    got, ref = _run(2300, 38, spec, dtype, G, 7, 4)
    _check(got, ref, dtype, f"v4 G={G} {dtype}")


def test_v4_agrees_with_v3_on_identical_inputs():
    spec = [(0, 0, 0), (0, 1, 1)]
    g4, ref = _run(5000, 128, spec, torch.bfloat16, 0, 3, 4)
    g3, _ = _run(5000, 128, spec, torch.bfloat16, 0, 3, 3)
    _check(g4, ref, torch.bfloat16, "v4")
    _check(g3, ref, torch.bfloat16, "v3")
    # both round the same fp32 value to bf16 up to MUFU / summation-order noise: at most a few bf16 ulps apart
    assert np.abs(g4 - g3).max() <= 2e-3 + 2.0 ** -6 * np.abs(ref).max()


@pytest.mark.parametrize("L", [1, 17, 513, 2300])
@pytest.mark.parametrize("rev", [0, 1])
def test_v7_no_replay_variant_vs_boundary_restatement(L, rev):
    """variant 7 (packed token pairs, no replay pass) on the same checker as v4."""
    got, ref = _run(L, 64, [(0, 0, rev)], torch.bfloat16, 0, 200 + L, 7)
    _check(got, ref, torch.bfloat16, f"v7 L={L} rev={rev}")


def test_v7_state_outputs_match_v3_fp32():
    """fp32 I/O, carry-in, end state, sum(dt) and the saved chunk states (the training / sharding hooks) must agree
    between the replay form (3) and the no-replay form (7): same recurrence, different association of the carry term."""
    from caduceus_b200 import functional as CF
    L, E = 1700, 48
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, [(0, 0, 0), (0, 1, 1)], torch.float32, 5)
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    h0 = torch.randn(2, E, 16, device=DEV)
    outs = {}
    for v in (3, 7):
        outs[v] = CF.scan_fwd(d(xz), d(delta), d(bc), tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)),
                              tuple(d(t) for t in tabs), L, h0=h0, want_state=True, want_chunk_state=True, variant=v)
    for got, ref, what in zip(outs[7], outs[3], ("out", "hlast", "dtsum", "chunk_state")):
        got, ref = got.float().cpu(), ref.float().cpu()
        if what == "out":
            got, ref = got[..., :L], ref[..., :L]
        assert torch.allclose(got, ref, rtol=2e-4, atol=2e-4 * float(ref.abs().max())), (what, (got - ref).abs().max())


# Opt-in cases (CAD_RUN_UNMEASURED=1).  Forward variants 9..12, backward variant 2 and conv_xproj's optional outputs passed the
# CPU emulation of their source AND a comparison with the default kernels on a B200 through the C-ABI (scripts/hw_probe.cu,
# profiles/r1_hw_probe_all_variants.log) — but with the last seconds of the round's GPU budget, so THESE pytest cases (ragged
# lengths, hooks, model level) have not run on hardware yet.  Variants 20..23 are emulation-verified only.  DESIGN.md §10.
unmeasured = pytest.mark.skipif(os.environ.get("CAD_RUN_UNMEASURED") != "1",
                                reason="pytest case not yet run on hardware (kernels: emulation + hw_probe); set "
                                       "CAD_RUN_UNMEASURED=1 to run it")


@unmeasured
@pytest.mark.parametrize("variant", [9, 10, 11, 12])
@pytest.mark.parametrize("L", [1, 17, 513, 2300])
@pytest.mark.parametrize("rev", [0, 1])
def test_v9_to_v12_vs_boundary_restatement(L, rev, variant):
    got, ref = _run(L, 64, [(0, 0, rev)], torch.bfloat16, 0, 300 + L, variant)
    _check(got, ref, torch.bfloat16, f"v{variant} L={L} rev={rev}")


@unmeasured
@pytest.mark.parametrize("variant", [9, 10, 11, 12])
@pytest.mark.parametrize("rev", [0, 1])
def test_v9_to_v12_hooks_vs_boundary_restatement(rev, variant):
    """conv halo + carry-in; end state, sum dt and saved chunk states; then the state-only pass."""
    from caduceus_b200 import functional as CF
    L, E, dtype = 1700, 40, torch.float16
    spec = [(0, 0, rev)]
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, 9)
    bc = bc.to(dtype).float()
    g = torch.Generator().manual_seed(1)
    halo, h0 = torch.randn(1, E, 3, generator=g).to(dtype), torch.randn(1, E, 16, generator=g)
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    args = (d(xz), d(delta), d(bc), tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)), tuple(d(t) for t in tabs), L)
    out, hlast, dtsum, cstate = CF.scan_fwd(*args, halo=d(halo), h0=d(h0), want_state=True, want_chunk_state=True,
                                            variant=variant)
    _, hl2, ds2, _ = CF.scan_fwd(*args, halo=d(halo), h0=d(h0), state_only=True, variant=variant)
    f = lambda t: t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk), [0], [0], [rev], L,
                       halo=f(halo), h0=f(h0), full=True)
    _check(out[..., :L].float().cpu().numpy(), ref[0], dtype, "out")
    for name, got, r in (("hlast", hlast, ref[1]), ("dtsum", dtsum, ref[2]), ("chunk_state", cstate, ref[3]),
                         ("hlast (state-only)", hl2, ref[1]), ("dtsum (state-only)", ds2, ref[2])):
        got = got.float().cpu().numpy()
        assert np.allclose(got, r, rtol=2e-3, atol=2e-3 * max(1.0, np.abs(r).max())), (name, np.abs(got - r).max())


@unmeasured
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


@unmeasured
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


@unmeasured
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


def test_v4_rejects_what_it_does_not_cover():
    from caduceus_b200 import functional as CF
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(100, 8, [(0, 0, 0)], torch.float32, 0)
    d = lambda t: t.to(DEV).contiguous()   # noqa: E731
    with pytest.raises(RuntimeError, match="variant 4"):
        CF.scan_fwd(d(xz), d(delta), d(bc), tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)),
                    tuple(d(t) for t in tabs), 100, variant=4)


@unmeasured
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_xproj_dt_epilogue_matches_softplus_of_its_own_dt_raw(dtype):
    """cad_conv_xproj_args.dt_b: the delta rows become dt = softplus(dt_raw + b) as fp16 bits; dt_raw itself (rounded to
    the io dtype, the reference's rounding point) is what the same kernel writes without dt_b; B / C rows are unchanged."""
    from caduceus_b200 import functional as CF
    L, E, R, N = 1500, 128, 8, 16
    g = torch.Generator().manual_seed(5)
    xz = torch.randn(2, 2 * E, 1504, generator=g).to(dtype).to(DEV)
    w_x = (torch.randn(2, R + 2 * N, E, generator=g) * E ** -0.5).to(dtype).to(DEV)
    w_dt = (torch.randn(2, E, R, generator=g) * R ** -0.5).to(dtype).to(DEV)
    conv_w4 = (0.5 * torch.randn(2, E, 4, generator=g)).to(DEV)
    conv_b = (0.1 * torch.randn(2, E, generator=g)).to(DEV)
    dt_b = (torch.randn(2, E, generator=g) * 2 - 3).to(DEV)
    dt_b[:, 0] = 25.0
    jobs = tuple(torch.tensor(v, dtype=torch.int32, device=DEV) for v in ([0, 0, 1, 1], [0, 1, 0, 1], [0, 1, 1, 0]))
    raw, bc0 = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L)
    dt16, bc1 = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, dt_b=dt_b)
    assert torch.equal(bc0, bc1)
    _, bc2, bcT = CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, want_bcT=True)     # token-major copy, variants 20..23
    assert torch.equal(bc0, bc2) and bcT.shape == (4, 1536, 2 * N)
    assert torch.equal(bcT[:, :L], bc0[..., :L].transpose(1, 2)) and (bcT[:, L:] == 0).all()
    got = (dt16 if dtype == torch.float16 else dt16.view(torch.float16))[..., :L].float()
    want = torch.nn.functional.softplus(raw[..., :L].float() + dt_b[jobs[1].long()][:, :, None])
    assert torch.isfinite(got).all()
    assert torch.allclose(got, want, rtol=2e-3, atol=1e-6), (got - want).abs().max()      # 2 fp16 ulps (MUFU ex2 / lg2)


@pytest.mark.parametrize("variant", [4, pytest.param(9, marks=unmeasured), pytest.param(10, marks=unmeasured),
                                     pytest.param(12, marks=unmeasured), pytest.param(-10, marks=unmeasured),
                                     pytest.param(-12, marks=unmeasured), pytest.param(20, marks=unmeasured)])
@pytest.mark.parametrize("tag", ["ps_small", "ph_config0"])
def test_model_forward_with_scan_variant_vs_reference_fixture(tag, variant):
    """The whole model with the scan forced to a non-default variant (9 / 10 take their 16-bit tile straight from the
    conv_xproj kernel), against the logits the reference's own code produced."""
    import caduceus
    from caduceus_b200 import functional as CF
    fx = golden(f"model_{tag}.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = caduceus.CaduceusForMaskedLM(cfg)
    model.load_state_dict(fx["state_dict"])
    model = model.to(DEV).to(torch.bfloat16).eval()
    launches = []
    orig = CF.scan_variant
    dt_in_xproj, variant = variant < 0, abs(variant)      # negative: softplus moved into conv_xproj (CAD_DT_IN_XPROJ)
    try:
        CF.SCAN_DT_IN_XPROJ = dt_in_xproj
        CF.SCAN_VARIANT = variant
        CF.scan_variant = lambda a: launches.append(orig(a)) or launches[-1]
        with torch.no_grad():
            logits = model(fx["input_ids"].to(DEV)).logits.float().cpu()
    finally:
        CF.SCAN_VARIANT = 0
        CF.SCAN_DT_IN_XPROJ = False
        CF.scan_variant = orig
    assert launches and all(v == variant for v in launches), launches
    # same criterion as test_gpu_parity.py::test_model_low_precision_vs_reference_fixture (16-bit stack vs fp32 fixture)
    rtol, atol = tol(torch.bfloat16)
    scale = fx["logits"].abs().max().item()
    err = (logits - fx["logits"]).abs().max().item()
    assert err <= atol + rtol * scale * 4, (err, scale)


# ---- backward variant 2 (csrc/scan_bwd_v2.cuh; CPU emulation: tests/test_emu_scan_bwd_v2.py) ---------------------------
def _bwd(L, E, spec, dtype, G, seed, variant, hooks=False):
    from caduceus_b200 import functional as CF
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    njobs, N = len(spec), 16
    g = torch.Generator().manual_seed(seed + 1)
    dout = torch.randn(njobs, E, ld, generator=g).to(dtype)
    halo = h0 = dhlast = None
    if hooks:
        halo = torch.randn(njobs, E, 3, generator=g).to(dtype)
        h0 = torch.randn(njobs, E, N, generator=g)
        dhlast = torch.randn(njobs, E, N, generator=g)
    seq, pset, rev = ([s_[k] for s_ in spec] for k in range(3))
    ref = boundary_grads(xz.float(), delta.float(), bc, dout.float(), conv_w4, conv_b, dt_b, A2, Dk, seq, pset, rev, L,
                         halo=None if halo is None else halo.float(), h0=h0, dhlast=dhlast)
    d = lambda t: None if t is None else t.to(DEV).contiguous()   # noqa: E731
    packed, jobs = tuple(d(t) for t in (conv_w4, conv_b, dt_b, A2, Dk)), tuple(d(t) for t in tabs)
    # the saved tensor comes from the forward KERNEL, as in training
    _, _, _, cstate = CF.scan_fwd(d(xz), d(delta), d(bc), packed, jobs, L, halo=d(halo), h0=d(h0), want_state=True,
                                  want_chunk_state=True)
    got = CF.scan_bwd(d(xz), d(delta), d(bc), d(dout), packed, jobs, L, cstate, halo=d(halo), h0=d(h0),
                      want_dh0=hooks, dhlast=d(dhlast), channels_per_cta=G, variant=variant)
    torch.cuda.synchronize()
    names = ("dz", "du", "ddelta", "dbc", "ddt_b", "dA2", "dD", "dh0")
    return {k: (None if v is None else v.float().cpu()) for k, v in zip(names, got)}, ref


def _check_grads(got, ref, L, dtype, what, hooks=False):
    eps = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 0.0}[dtype]
    for k in ("dz", "du", "ddelta", "dbc", "ddt_b", "dA2", "dD") + (("dh0",) if hooks else ()):
        g_, r = got[k].double(), ref[k].double()
        if k in ("dz", "du", "ddelta", "dbc"):
            g_ = g_[..., :L]
        assert torch.isfinite(g_).all(), (what, k)
        scale = max(1.0, float(r.abs().max()))
        err, bound = (g_ - r).abs(), 2e-3 * scale + (2 * eps + 2e-3) * r.abs()     # MUFU approximations: 2e-3 relative
        assert torch.all(err <= bound), f"{what} {k}: max err {err.max():.3e} (scale {scale:.3e})"


@unmeasured
@pytest.mark.parametrize("L", [1, 9, 255, 513, 1030, 2300])
@pytest.mark.parametrize("rev", [0, 1])
def test_bwd_v2_vs_float64_autograd_at_the_boundary(L, rev):
    got, ref = _bwd(L, 24, [(0, 0, rev)], torch.float32, 0, 400 + L, 2)
    _check_grads(got, ref, L, torch.float32, f"bwd v2 L={L} rev={rev}")


@unmeasured
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("G", [1, 3, 7])
def test_bwd_v2_ps_job_layout_hooks_and_cta_shapes(dtype, G):
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    got, ref = _bwd(1700, 19, spec, dtype, G, 17, 2, hooks=True)
    _check_grads(got, ref, 1700, dtype, f"bwd v2 G={G} {dtype}", hooks=True)


@unmeasured
def test_bwd_v2_agrees_with_v1_on_identical_inputs():
    spec = [(0, 0, 0), (0, 1, 1)]
    g2, ref = _bwd(5000, 64, spec, torch.float32, 0, 3, 2)
    g1, _ = _bwd(5000, 64, spec, torch.float32, 0, 3, 1)
    _check_grads(g1, ref, 5000, torch.float32, "bwd v1")
    _check_grads(g2, ref, 5000, torch.float32, "bwd v2")


@unmeasured
def test_mixer_backward_with_bwd_v2_vs_oracle_autograd():
    """BiMambaWrapper.backward with the scan gradient forced to variant 2, against autograd through the CPU oracle."""
    import caduceus_b200
    from caduceus_b200 import functional as CF
    from test_gpu_backward import _grad_close, _oracle_mixer_grads
    fx = golden("mixer_add_tied.pt")
    m = caduceus_b200.BiMambaWrapper(fx["d_model"], bidirectional=True, bidirectional_strategy=fx["strategy"],
                                     bidirectional_weight_tie=fx["tie"], **fx["ssm_cfg"])
    m.load_state_dict(fx["state_dict"])
    m = m.to(DEV)
    torch.manual_seed(0)
    h, gout = torch.randn(2, 1537, fx["d_model"]), torch.randn(2, 1537, fx["d_model"])
    hd = h.to(DEV).requires_grad_(True)
    try:
        CF.SCAN_BWD_VARIANT = 2
        m(hd).backward(gout.to(DEV))
    finally:
        CF.SCAN_BWD_VARIANT = 0
    _, ref_dh, ref_dp = _oracle_mixer_grads(fx["state_dict"], h, gout, fx["strategy"])
    _grad_close(hd.grad, ref_dh, 5e-3, 1e-3, "d hidden")
    for name, p in m.named_parameters():
        if ref_dp[name] is not None:
            _grad_close(p.grad, ref_dp[name].reshape(p.shape), 5e-3, 2e-3, f"d {name}")
