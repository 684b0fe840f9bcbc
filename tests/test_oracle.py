"""CPU tests of the ORACLE: known answers, an independent implementation, the reference-generated fixtures and the
reference's own RC-equivariance properties (ref:caduceus/tests/test_rcps.py) — so the checker itself is pinned."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ORACLE, golden, tol
from mamba_ssm.ops.selective_scan_interface import causal_conv1d_ref, mamba_inner_ref, selective_scan_ref
from mamba_ssm.ops.triton.layernorm import RMSNorm, layer_norm_fn, rms_norm_fn
import caduceus_oracle as CO

MODEL_FIXTURES = ["ph_config0", "ps_config0", "ph_small", "ps_small", "ps_nonfused", "ph_nonfused_ln",
                  "ps_ln_fp32res", "ph_mul_untied", "ph_unidir", "ps_d118", "ph_d118"]


# ---- closed-form known answers for the selective scan ----------------------------
def test_scan_geometric_series():
    L, d, n = 50, 3, 4
    delta = torch.full((1, d, L), 0.3)
    u = torch.full((1, d, L), 2.0)
    A = -torch.arange(1, n + 1, dtype=torch.float32).repeat(d, 1)
    B = torch.ones(1, n, L)
    C = torch.ones(1, n, L)
    y = selective_scan_ref(u, delta, A, B, C)
    a = torch.exp(0.3 * A[0])                       # per-state decay
    t = torch.arange(1, L + 1, dtype=torch.float32)
    expect = ((1 - a[None, :] ** t[:, None]) / (1 - a[None, :]) * 0.6).sum(-1)     # b = delta*B*u = 0.6
    assert torch.allclose(y[0, 0], expect, rtol=1e-5, atol=1e-5)


def test_scan_impulse_decay_and_skip():
    L, n = 40, 2
    u = torch.zeros(1, 1, L); u[0, 0, 5] = 1.0
    delta = torch.full((1, 1, L), 0.5)
    A = torch.tensor([[-1.0, -2.0]])
    B = torch.ones(1, n, L); C = torch.ones(1, n, L)
    y = selective_scan_ref(u, delta, A, B, C, D=torch.tensor([3.0]))
    k = torch.arange(0, L - 5, dtype=torch.float32)
    expect = 0.5 * (torch.exp(-0.5 * k) + torch.exp(-1.0 * k))
    expect[0] += 3.0                                  # D*u at the impulse
    assert torch.allclose(y[0, 0, 5:], expect, rtol=1e-5, atol=1e-6)
    assert torch.all(y[0, 0, :5] == 0)


def test_scan_delta_to_zero_is_skip_and_len1():
    u = torch.randn(2, 4, 7)
    y = selective_scan_ref(u, torch.full_like(u, -40.0), -torch.ones(4, 3), torch.randn(2, 3, 7), torch.randn(2, 3, 7),
                           D=torch.arange(4.0), delta_softplus=True)
    assert torch.allclose(y, u * torch.arange(4.0)[None, :, None], atol=1e-6)
    u1 = torch.randn(1, 2, 1)
    y1, last = selective_scan_ref(u1, torch.ones(1, 2, 1), -torch.ones(2, 3), torch.ones(1, 3, 1), torch.ones(1, 3, 1),
                                  return_last_state=True)
    assert torch.allclose(y1[..., 0], 3 * u1[..., 0]) and last.shape == (1, 2, 3)


def test_scan_fixtures_reproduce():
    for name in ("scan_upstream_ranges.pt", "scan_init_ranges.pt"):
        fx = golden(name)
        out, last = selective_scan_ref(fx["u"], fx["delta"], fx["A"], fx["B"], fx["C"], fx["D"], fx["z"],
                                       fx["delta_bias"], delta_softplus=True, return_last_state=True)
        assert torch.equal(out, fx["out"]) and torch.equal(last, fx["last_state"])


# ---- independent implementation: transformers' MambaMixer.slow_forward ----------------------------------------
def test_mamba_matches_transformers_slow_forward():
    from transformers.models.mamba.configuration_mamba import MambaConfig
    from transformers.models.mamba.modeling_mamba import MambaMixer
    from mamba_ssm.modules.mamba_simple import Mamba
    torch.manual_seed(0)
    d_model = 32
    hf_cfg = MambaConfig(hidden_size=d_model, state_size=16, conv_kernel=4, expand=2, time_step_rank=2,
                         use_bias=False, use_conv_bias=True, num_hidden_layers=1, vocab_size=16)
    hf = MambaMixer(hf_cfg, layer_idx=0).eval()
    ours = Mamba(d_model, dt_rank=2).eval()
    with torch.no_grad():
        for p in hf.parameters():
            if p.dim() > 1:
                p.normal_(std=0.2)
        hf.A_log.copy_(torch.log(torch.rand_like(hf.A_log) * 4 + 0.5))
        hf.dt_proj.bias.normal_(mean=-2.0)
        hf.D.normal_()
        ours.load_state_dict(hf.state_dict())
    h = torch.randn(2, 37, d_model)
    with torch.no_grad():
        ref = hf.slow_forward(h)
        got = ours(h)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), (got - ref).abs().max()


# ---- norm restatement ---------------------------------------------------------------------------------------------
def test_norm_restatement():
    torch.manual_seed(1)
    x, r = torch.randn(3, 5, 16), torch.randn(3, 5, 16)
    w, b = torch.randn(16), torch.randn(16)
    y, res = layer_norm_fn(x, w, b, residual=r, eps=1e-5, prenorm=True)
    assert torch.allclose(res, x + r) and torch.allclose(y, F.layer_norm(x + r, (16,), w, b, 1e-5), atol=1e-5)
    y2 = rms_norm_fn(x, w, None, eps=1e-5)
    assert torch.allclose(y2, x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w, atol=1e-6)
    # upstream: a residual that is passed in keeps its dtype; residual_in_fp32 decides the dtype of a NEW residual stream only
    yb, resb = rms_norm_fn(x.bfloat16(), w.bfloat16(), None, residual=r.bfloat16(), prenorm=True, residual_in_fp32=True)
    assert yb.dtype == torch.bfloat16 and resb.dtype == torch.bfloat16
    yb, resb = rms_norm_fn(x.bfloat16(), w.bfloat16(), None, residual=r, prenorm=True, residual_in_fp32=False)
    assert yb.dtype == torch.bfloat16 and resb.dtype == torch.float32
    yb, resb = rms_norm_fn(x.bfloat16(), w.bfloat16(), None, residual=None, prenorm=True, residual_in_fp32=True)
    assert yb.dtype == torch.bfloat16 and resb.dtype == torch.float32


# ---- model-level restatement vs fixtures produced by the reference's own code ------------------------------------
@pytest.mark.parametrize("tag", MODEL_FIXTURES)
def test_model_restatement_matches_reference_fixture(tag):
    fx = golden(f"model_{tag}.pt")
    if fx["input_ids"].shape[1] > 512:
        ids = fx["input_ids"]
    else:
        ids = fx["input_ids"]
    with torch.no_grad():
        logits, hidden = CO.model_ref(ids, fx["state_dict"], fx["config"], return_hidden=True)
    assert torch.allclose(logits, fx["logits"], rtol=1e-5, atol=1e-5), (logits - fx["logits"]).abs().max()
    assert torch.allclose(hidden, fx["last_hidden_state"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["add_tied", "mul_untied", "len1"])
def test_mixer_restatement_matches_reference_fixture(tag):
    fx = golden(f"mixer_{tag}.pt")
    with torch.no_grad():
        out = CO.bimamba_ref(fx["hidden"], fx["state_dict"], "", True, fx["strategy"])
    assert torch.allclose(out, fx["out"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("rcps", [False, True])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_block_restatement_matches_reference_fixture(rcps, fused, with_res):
    tag = f"{'ps' if rcps else 'ph'}_{'fused' if fused else 'plain'}_{'res' if with_res else 'nores'}"
    fx = golden(f"block_{tag}.pt")
    cfg = dict(rcps=rcps, fused_add_norm=fused, rms_norm=True, norm_epsilon=1e-5, residual_in_fp32=False,
               bidirectional=True, bidirectional_strategy="add")
    sd = {"L." + k: v for k, v in fx["state_dict"].items()}
    with torch.no_grad():
        h, r = CO.block_ref(fx["hidden"], fx["residual"], sd, "L.", cfg)
    assert torch.allclose(h, fx["out_hidden"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(r, fx["out_residual"], rtol=1e-5, atol=1e-6)


# ---- the reference's RC-equivariance property (ref:caduceus/tests/test_rcps.py:341-419) on the oracle ----------
@pytest.mark.parametrize("tag", ["ps_small", "ps_nonfused"])
def test_rc_equivariance_of_oracle_logits(tag):
    fx = golden(f"model_{tag}.pt")
    cfg, sd = fx["config"], fx["state_dict"]
    _, cmap = CO.padded_cmap(cfg)
    ids = fx["input_ids"]
    rc_ids = cmap[torch.flip(ids, dims=[-1])]
    with torch.no_grad():
        out = CO.model_ref(ids, sd, cfg)
        out_rc = CO.model_ref(rc_ids, sd, cfg)
    assert torch.allclose(out, torch.flip(out_rc[..., cmap], dims=[1]), rtol=6e-4, atol=2e-3)


# ---- the reference's own modules, live, when /root/reference is present (build container only) -------------------
def test_live_reference_agrees_with_fixture():
    from ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("/root/reference not present (GPU box): fixtures stand in")
    ref = load_reference()
    fx = golden("model_ps_small.pt")
    cfg = ref.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = ref.CaduceusForMaskedLM(cfg).eval()
    model.load_state_dict(fx["state_dict"])
    with torch.no_grad():
        logits = model(fx["input_ids"]).logits
    assert torch.equal(logits, fx["logits"])


# ---- C restatement vs torch restatement -------------------------------------------------------------------------------
def _load_cscan():
    path = os.path.join(ORACLE, "libcscan.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.run(["make", "-C", ORACLE], check=True, capture_output=True)
    lib = ctypes.CDLL(path)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.cad_oracle_mamba_inner.restype = ctypes.c_int
    lib.cad_oracle_mamba_inner.argtypes = [fp] * 8 + [ctypes.c_long] * 5 + [ctypes.c_int, fp, fp]
    return lib


def c_mamba_inner(xz, conv_w, conv_b, w_x, w_dt, dt_b, A, D, rev):
    lib = _load_cscan()
    twoE, L = xz.shape
    E = twoE // 2
    N, R, K = A.shape[1], w_dt.shape[1], conv_w.shape[1]
    arrs = [np.ascontiguousarray(t.numpy(), dtype=np.float32) for t in (xz, conv_w, conv_b, w_x, w_dt, dt_b, A, D)]
    y = np.empty((E, L), dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib.cad_oracle_mamba_inner(*[a.ctypes.data_as(fp) for a in arrs], L, E, N, R, K, int(rev),
                                    y.ctypes.data_as(fp), None)
    assert rc == 0
    return torch.from_numpy(y)


@pytest.mark.parametrize("rev", [0, 1])
def test_c_restatement_matches_torch(rev):
    torch.manual_seed(3)
    E, N, R, K, L = 24, 16, 3, 4, 301
    xz = torch.randn(2 * E, L)
    conv_w, conv_b = torch.randn(E, K) * 0.5, torch.randn(E) * 0.1
    w_x, w_dt = torch.randn(R + 2 * N, E) * 0.2, torch.randn(E, R) * 0.5
    dt_b, A, D = torch.randn(E) - 3.0, -torch.rand(E, N) * 4 - 0.1, torch.randn(E)
    got = c_mamba_inner(xz, conv_w, conv_b, w_x, w_dt, dt_b, A, D, rev)
    xin = xz.flip(-1) if rev else xz
    eye = torch.eye(E)
    ref = mamba_inner_ref(xin[None], conv_w[:, None, :], conv_b, w_x, w_dt, eye, None, A, None, None, D,
                          delta_bias=dt_b, delta_softplus=True)[0].t()
    if rev:
        ref = ref.flip(-1)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5), (got - ref).abs().max()
