"""GPU gradient parity (-m gpu): backward kernels through the C-ABI against torch autograd of the CPU oracle
(oracle/caduceus_oracle.py, fp64 or fp32) on identical seeded inputs and weights."""
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _grad_close(got, ref, rtol, atol, what):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    scale = ref.abs().max().clamp_min(1e-30)
    bound = atol * scale + rtol * ref.abs()
    assert torch.all(err <= bound), f"{what}: max err {err.max():.3e} (ref max {scale:.3e}), excess {(err - bound).max():.3e}"


def _oracle_mixer_grads(sd, h, gout, strategy, dtype=torch.float64):
    import caduceus_oracle as CO
    sd64 = {k: v.detach().to(dtype).requires_grad_(True) for k, v in sd.items()}
    # tied projections share one tensor in the oracle too
    for proj in ("in_proj.weight", "out_proj.weight"):
        if torch.equal(sd["mamba_fwd." + proj], sd["mamba_rev." + proj]):
            sd64["mamba_rev." + proj] = sd64["mamba_fwd." + proj]
    h64 = h.detach().to(dtype).requires_grad_(True)
    out = CO.bimamba_ref(h64, sd64, "", True, strategy)
    out.backward(gout.to(dtype))
    return out.detach(), h64.grad, {k: v.grad for k, v in sd64.items()}


@pytest.mark.parametrize("tag,L", [("add_tied", 700), ("mul_untied", 130), ("add_tied", 16), ("add_tied", 1537)])
def test_mixer_backward_vs_oracle_autograd(tag, L):
    import caduceus_b200
    fx = golden(f"mixer_{tag}.pt")
    torch.manual_seed(0)
    d = fx["d_model"]
    m = caduceus_b200.BiMambaWrapper(d, bidirectional=True, bidirectional_strategy=fx["strategy"],
                                     bidirectional_weight_tie=fx["tie"], **fx["ssm_cfg"])
    m.load_state_dict(fx["state_dict"])
    m = m.to(DEV)
    h = torch.randn(2, L, d)
    gout = torch.randn(2, L, d)
    hd = h.to(DEV).requires_grad_(True)
    out = m(hd)
    out.backward(gout.to(DEV))
    ref_out, ref_dh, ref_dp = _oracle_mixer_grads(fx["state_dict"], h, gout, fx["strategy"])
    _grad_close(out.detach(), ref_out, 2e-3, 2e-4, "forward (training path)")
    _grad_close(hd.grad, ref_dh, 5e-3, 1e-3, "d hidden")
    for name, p in m.named_parameters():
        ref = ref_dp[name]
        if ref is None:          # tied duplicate: gradient lives on the shared tensor
            continue
        _grad_close(p.grad, ref.reshape(p.shape), 5e-3, 2e-3, f"d {name}")


@pytest.mark.parametrize("tag", ["ps_small", "ph_small", "ps_nonfused", "ph_nonfused_ln", "ps_ln_fp32res", "ps_d118", "ph_d118"])
def test_model_loss_backward_vs_oracle_autograd(tag):
    """CaduceusForMaskedLM: MLM cross-entropy loss and ALL parameter gradients vs autograd through the oracle."""
    import caduceus
    import caduceus_oracle as CO
    fx = golden(f"model_{tag}.pt")
    cfgd = fx["config"]
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in cfgd.items()})
    cfg.pad_token_id = 4
    model = caduceus.CaduceusForMaskedLM(cfg)
    model.load_state_dict(fx["state_dict"])
    model = model.to(DEV).train()
    ids = fx["input_ids"]
    g = torch.Generator().manual_seed(0)
    labels = ids.clone()
    labels[torch.rand(ids.shape, generator=g) > 0.3] = 4          # ignore_index = PAD, ref:configs/experiment/hg38/hg38.yaml:9-11
    out = model(ids.to(DEV), labels=labels.to(DEV))
    out.loss.backward()

    sd = {k: v.detach().double().requires_grad_(True) for k, v in fx["state_dict"].items()}
    # re-tie what the model ties (shared storage in the reference's state_dict)
    names = list(sd)
    for k in names:
        if ".mamba_rev.in_proj.weight" in k or ".mamba_rev.out_proj.weight" in k:
            sd[k] = sd[k.replace("mamba_rev", "mamba_fwd")]
    emb_key = ("caduceus.backbone.embeddings.word_embeddings.embedding.weight" if cfgd["rcps"]
               else "caduceus.backbone.embeddings.word_embeddings.weight")
    head_key = "lm_head.lm_head.weight" if cfgd["rcps"] else "lm_head.weight"
    sd[head_key] = sd[emb_key]
    logits = CO.model_ref(ids, sd, cfgd)
    loss = torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1), ignore_index=4)
    loss.backward()
    assert abs(out.loss.item() - loss.item()) < 2e-4 * max(1.0, abs(loss.item()))
    checked = 0
    for name, p in model.named_parameters():
        ref = sd[name].grad
        if ref is None:
            continue
        _grad_close(p.grad, ref.reshape(p.shape), 2e-2, 5e-3, f"d {name}")
        checked += 1
    assert checked >= 10


def test_bf16_autocast_train_step_reduces_loss():
    """A few AdamW steps under bf16 autocast on a tiny PS model: finite grads, loss goes down."""
    import caduceus
    torch.manual_seed(0)
    cmap = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 10, 8: 9, 9: 8, 10: 7, 11: 11}
    cfg = caduceus.CaduceusConfig(d_model=64, n_layer=2, vocab_size=12, ssm_cfg={"d_state": 16}, rms_norm=True,
                                  residual_in_fp32=False, fused_add_norm=True, rcps=True, complement_map=cmap,
                                  pad_token_id=4)
    model = caduceus.CaduceusForMaskedLM(cfg).to(DEV).train()
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3)
    ids = torch.randint(7, 11, (4, 1024), device=DEV)
    labels = ids.clone()
    inp = ids.clone()
    mask = torch.rand(ids.shape, device=DEV) < 0.15
    inp[mask] = 3
    labels[~mask] = 4
    losses = []
    for _ in range(8):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = model(inp, labels=labels).loss
        opt.zero_grad(set_to_none=True)
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.05, losses


@pytest.mark.parametrize("rcps", [False, True])
@pytest.mark.parametrize("D,V", [(512, 16), (1024, 16), (118, 16), (64, 4096)])
def test_embedding_backward_any_width_and_vocabulary(D, V, rcps):
    """dW of the (RC-equivariant) embedding for widths beyond one 48 KB tile (d_model 512 / 1024 with the padded vocabulary of 16
    used to fail at the first backward), a width that is not a multiple of the vector size, and a large vocabulary — against
    autograd through the literal reference formula (ref:caduceus/modeling_rcps.py:54-67)."""
    from caduceus_b200 import functional as CF
    g = torch.Generator().manual_seed(D + V)
    W = torch.randn(V, D, generator=g).to("cuda").requires_grad_(True)
    ids = torch.randint(0, V, (3, 257), generator=g).to("cuda")
    cmap = torch.randperm(V, generator=g).to("cuda") if rcps else None
    out = CF.embedding(ids, W, cmap)
    gout = torch.randn(out.shape, generator=g).to("cuda")
    (dW,) = torch.autograd.grad(out, W, gout)
    Wr = W.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.embedding(ids, Wr)
    if rcps:
        rc = torch.flip(torch.nn.functional.embedding(cmap[torch.flip(ids, dims=[-1])], Wr), dims=[-2, -1])
        ref = torch.cat([ref, rc], dim=-1)
    assert torch.equal(out, ref)                      # index work: bit-exact
    (dWr,) = torch.autograd.grad(ref, Wr, gout)
    assert torch.allclose(dW, dWr, rtol=1e-5, atol=1e-4), (dW - dWr).abs().max()


@pytest.mark.parametrize("rcps", [False, True])
@pytest.mark.parametrize("dtype,weighted", [(torch.float32, False), (torch.bfloat16, False), (torch.float32, True)])
def test_fused_lm_head_cross_entropy_vs_reference_formula(rcps, dtype, weighted):
    """csrc/head_ce.cu against the literal reference: logits = head(hidden).float() (RCPS: x1 W^T +
    flip(x2) W[cmap]^T, ref:caduceus/modeling_rcps.py:233-246), then (weighted) cross-entropy with ignore_index
    (ref:caduceus/modeling_caduceus.py:279-294) — loss, d hidden (zero rows at ignored positions) and d weight."""
    from caduceus_b200 import functional as CF
    g = torch.Generator().manual_seed(3 + rcps)
    B, L, D, V = 2, 700, 96, 16
    width = 2 * D if rcps else D
    hid = torch.randn(B, L, width, generator=g).to(DEV).to(dtype).requires_grad_(True)
    W = (0.3 * torch.randn(V, D, generator=g)).to(DEV).requires_grad_(True)
    cmap = torch.tensor([0, 1, 2, 3, 4, 5, 6, 10, 9, 8, 7, 11, 12, 13, 14, 15], device=DEV) if rcps else None
    labels = torch.randint(0, 12, (B, L), generator=g).to(DEV)
    labels[torch.rand(B, L, generator=g).to(DEV) > 0.15] = 4                 # ~85 % ignored, as in MLM
    lw = torch.rand(B, L, generator=g).to(DEV) if weighted else None
    loss = CF.lm_head_cross_entropy(hid, W, labels, cmap=cmap, loss_weights=None if lw is None else lw.clone(), ignore_index=4)
    dh, dw = torch.autograd.grad(loss, (hid, W))

    hid_r = hid.detach().clone().requires_grad_(True)
    W_r = W.detach().clone().requires_grad_(True)
    Wc = W_r.to(dtype)
    logits = torch.nn.functional.linear(hid_r[..., :D], Wc)
    if rcps:
        logits = logits + torch.nn.functional.linear(torch.flip(hid_r[..., D:], dims=[-1]), Wc[cmap, :])
    logits = logits.float().view(-1, V)
    y = labels.view(-1)
    if weighted:
        ce = torch.nn.functional.cross_entropy(logits, y, ignore_index=4, reduction="none")
        w = lw.view(-1).clone()
        w[y == 4] = 0.0
        ref = (ce * (w / w.sum())).sum()
    else:
        ref = torch.nn.functional.cross_entropy(logits, y, ignore_index=4)
    dh_r, dw_r = torch.autograd.grad(ref, (hid_r, W_r))
    # bf16: the reference rounds its logits to bf16 before the softmax; here they stay fp32 — compare at bf16 resolution of the range
    rt, at = (1e-4, 1e-6) if dtype == torch.float32 else (2e-2, 2e-3)
    assert abs(loss.item() - ref.item()) <= rt * abs(ref.item()) + 1e-5, (loss.item(), ref.item())
    assert (dh.float()[labels == 4] == 0).all()
    _grad_close(dh.float(), dh_r.float(), rt * 10, at * 10, "d hidden")
    _grad_close(dw, dw_r, rt * 10, at * 10, "d weight")


@pytest.mark.parametrize("tag", ["ps_small", "ph_small"])
def test_model_fused_head_loss_matches_the_logits_path(tag):
    """config.fused_head_loss: same loss and parameter gradients as the reference-shaped path that materialises the logits."""
    import caduceus
    fx = golden(f"model_{tag}.pt")
    grads = {}
    for fused in (False, True):
        cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
        cfg.pad_token_id = 4
        cfg.fused_head_loss = fused
        model = caduceus.CaduceusForMaskedLM(cfg)
        model.load_state_dict(fx["state_dict"])
        model = model.to(DEV).train()
        ids = fx["input_ids"]
        g = torch.Generator().manual_seed(0)
        labels = ids.clone()
        labels[torch.rand(ids.shape, generator=g) > 0.3] = 4
        out = model(ids.to(DEV), labels=labels.to(DEV))
        assert (out.logits is None) == fused
        out.loss.backward()
        grads[fused] = (out.loss.item(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None})
    assert abs(grads[True][0] - grads[False][0]) < 1e-5 * max(1.0, abs(grads[False][0]))
    assert set(grads[True][1]) == set(grads[False][1])
    for n, gr in grads[False][1].items():
        _grad_close(grads[True][1][n], gr, 1e-3, 1e-5, f"d {n}")
