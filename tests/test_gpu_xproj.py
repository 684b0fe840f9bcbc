"""GPU parity of the tcgen05 / TMEM conv + SiLU -> x_proj -> dt_proj kernel (cad_conv_xproj_fwd, csrc/xproj.cu) through the
C-ABI: against a float64 restatement of upstream's pipeline at the kernel boundary (causal_conv1d -> x_proj -> dt_proj with
the reference's rounding points) over ragged lengths, channel counts, dt ranks, directions and the shard
halo, and end to end through the model against the fixtures produced by the reference's own code.  (The oracle's own
intermediates are compared in tests/test_gpu_parity.py::test_conv_xproj_outputs_vs_oracle_intermediates.)"""
import pytest
import torch

from conftest import golden, tol

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _problem(L, E, R, dtype, spec, seed, nseq=2, halo=False):
    N = 16
    g = torch.Generator().manual_seed(seed)
    ld = (max(L, 1) + 15) // 16 * 16
    xz = torch.randn(nseq, 2 * E, ld, generator=g).to(dtype)
    xz[..., L:] = 0
    P = 2
    w_x = (torch.randn(P, R + 2 * N, E, generator=g) * E ** -0.5).to(dtype)
    w_dt = (torch.randn(P, E, R, generator=g) * R ** -0.5).to(dtype)
    conv_w4 = 0.5 * torch.randn(P, E, 4, generator=g)
    conv_b = 0.1 * torch.randn(P, E, generator=g)
    tabs = tuple(torch.tensor([s[i] for s in spec], dtype=torch.int32) for i in range(3))
    h = torch.randn(len(spec), E, 3, generator=g).to(dtype) if halo else None
    return xz, w_x, w_dt, conv_w4, conv_b, tabs, h


def _restatement(xz, w_x, w_dt, conv_w4, conv_b, spec, L, dtype, halo=None):
    """float64 at the kernel boundary: u = silu(conv(x) + b) rounded to the io dtype, x_dbl = W_x u (B / C rows kept in full
    precision), dt_raw = W_dt . round(x_dbl[:R]).  Reversed jobs: anti-causal taps (logical time runs right to left)."""
    E = xz.shape[1] // 2
    R = w_dt.shape[-1]
    deltas, bcs = [], []
    for j, (s, p, rev) in enumerate(spec):
        x = xz[s, :E, :L].double()
        if rev:
            x = x.flip(-1)
        pre = halo[j].double() if halo is not None else torch.zeros(E, 3, dtype=torch.float64)
        xp = torch.cat([pre, x], dim=-1)                                  # logical times -3 .. L-1
        w = conv_w4[p].double()
        acc = conv_b[p].double()[:, None] + sum(w[:, k:k + 1] * xp[:, k:k + L] for k in range(4))
        u = (acc * torch.sigmoid(acc)).to(dtype).double()
        x_dbl = w_x[p].double() @ u                                       # (R + 2N, L)
        dt_raw = w_dt[p].double() @ x_dbl[:R].to(dtype).double()
        bc = x_dbl[R:]
        if rev:
            dt_raw, bc = dt_raw.flip(-1), bc.flip(-1)
        deltas.append(dt_raw)
        bcs.append(bc)
    return torch.stack(deltas), torch.stack(bcs)


def _run(xz, w_x, w_dt, conv_w4, conv_b, tabs, L, halo, want_bcT=True, packed=False):
    from caduceus_b200 import functional as CF
    d = lambda t: None if t is None else t.to(DEV).contiguous()   # noqa: E731
    wxp = CF.pack_w_x(d(w_x), w_dt.shape[-1]) if packed else None
    out = CF.conv_xproj(d(xz), d(w_x), d(w_dt), d(conv_w4), d(conv_b), tuple(d(t) for t in tabs), L, halo=d(halo),
                        want_bcT=want_bcT, w_x_packed=wxp)
    torch.cuda.synchronize()
    return out


SPEC4 = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("L,E,R", [(1, 64, 4), (7, 128, 8), (127, 128, 8), (128, 512, 16), (129, 192, 12), (1500, 256, 16),
                                    (3000, 512, 16), (40001, 512, 16), (5000, 1024, 16)])
def test_xproj_vs_restatement(L, E, R, dtype, packed):
    """Ragged lengths around the 128-token tile, d_inner with a 64-channel tail chunk (192), dt_rank < 16 (zero-padded operand
    rows), four jobs mixing sequences, parameter sets and directions — more tiles than persistent CTAs at L = 40001."""
    args = _problem(L, E, R, dtype, SPEC4, 100 + L)
    xz, w_x, w_dt, conv_w4, conv_b, tabs, _ = args
    delta, bc, bcT = _run(xz, w_x, w_dt, conv_w4, conv_b, tabs, L, None, packed=packed)   # packed: W_x handed over pre-arranged per slab
    want_delta, want_bc = _restatement(xz, w_x, w_dt, conv_w4, conv_b, SPEC4, L, dtype)
    rtol, atol = tol(dtype)
    got_bc, got_delta = bc[..., :L].double().cpu(), delta[..., :L].double().cpu()
    assert torch.isfinite(got_bc).all() and torch.isfinite(got_delta).all()
    # B / C: fp32 accumulation of products of 16-bit values; u itself may differ from the restatement by one rounding of the
    # activation (tanh.approx SiLU) on a few elements
    assert torch.all((got_bc - want_bc).abs() <= 0.1 * atol + 0.1 * rtol * want_bc.abs()), (got_bc - want_bc).abs().max()
    assert torch.all((got_delta - want_delta).abs() <= 0.5 * atol + rtol * want_delta.abs()), (got_delta - want_delta).abs().max()
    # padding contract: zeros beyond L in bc (every column < ldbc) and bcT (rows < ceil128(L)); bcT == bc transposed
    assert (bc[..., L:] == 0).all() and (bcT[:, L:] == 0).all()
    assert torch.equal(bcT[:, :L], bc[..., :L].transpose(1, 2))


@pytest.mark.parametrize("rev", [0, 1])
def test_xproj_conv_halo_of_a_sequence_shard(rev):
    """Shard hook: the three x samples that logically precede the shard enter the conv of the first / last tile."""
    L, E, R, dtype = 700, 128, 8, torch.bfloat16
    spec = [(0, 0, rev), (1, 1, 1 - rev)]
    xz, w_x, w_dt, conv_w4, conv_b, tabs, halo = _problem(L, E, R, dtype, spec, 7 + rev, halo=True)
    delta, bc = _run(xz, w_x, w_dt, conv_w4, conv_b, tabs, L, halo, want_bcT=False)
    want_delta, want_bc = _restatement(xz, w_x, w_dt, conv_w4, conv_b, spec, L, dtype, halo=halo)
    rtol, atol = tol(dtype)
    got_bc, got_delta = bc[..., :L].double().cpu(), delta[..., :L].double().cpu()
    assert torch.all((got_bc - want_bc).abs() <= 0.1 * atol + 0.1 * rtol * want_bc.abs()), (got_bc - want_bc).abs().max()
    assert torch.all((got_delta - want_delta).abs() <= 0.5 * atol + rtol * want_delta.abs()), (got_delta - want_delta).abs().max()


def test_xproj_is_deterministic_and_reentrant():
    """Back-to-back launches on one stream (TMEM allocated and released per CTA, two CTAs per SM) give identical bits."""
    L, E, R, dtype = 20000, 512, 16, torch.bfloat16
    xz, w_x, w_dt, conv_w4, conv_b, tabs, _ = _problem(L, E, R, dtype, SPEC4, 3)
    outs = [_run(xz, w_x, w_dt, conv_w4, conv_b, tabs, L, None) for _ in range(3)]
    for o in outs[1:]:
        assert all(torch.equal(a[..., :L] if a.dim() == 3 and a.shape[-1] >= L else a, b[..., :L] if b.dim() == 3 and b.shape[-1] >= L else b)
                   for a, b in zip(o[:2], outs[0][:2]))
        assert torch.equal(o[2], outs[0][2])


@pytest.mark.parametrize("packed", [True, False])
@pytest.mark.parametrize("tag", ["ps_small", "ph_config0"])
def test_model_forward_runs_the_tensor_core_projection_kernel(tag, packed):
    """The bf16 model forward goes through cad_conv_xproj_fwd (not the unfused conv + cuBLAS path) and matches the logits the
    reference's own code produced."""
    import caduceus
    from caduceus_b200 import functional as CF
    fx = golden(f"model_{tag}.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = caduceus.CaduceusForMaskedLM(cfg)
    model.load_state_dict(fx["state_dict"])
    model = model.to(DEV).to(torch.bfloat16).eval()
    calls = []
    lib_fn, prev = CF.conv_xproj, CF.XPROJ_PACK_W
    try:
        CF.XPROJ_PACK_W = packed
        CF.conv_xproj = lambda *a, **k: calls.append(k.get("w_x_packed") is not None) or lib_fn(*a, **k)
        with torch.no_grad():
            logits = model(fx["input_ids"].to(DEV)).logits.float().cpu()
    finally:
        CF.conv_xproj, CF.XPROJ_PACK_W = lib_fn, prev
    assert len(calls) == cfg.n_layer and all(c == packed for c in calls)
    ref = fx["logits"].float()
    rtol, atol = tol(torch.bfloat16)
    scale = ref.abs().max().item()
    assert (logits - ref).abs().max().item() <= atol + rtol * scale
