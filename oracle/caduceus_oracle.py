"""ORACLE (test infrastructure only) — CPU restatement of the reference's model-level algebra ABOVE the mamba_ssm
boundary, written as plain functions over a reference-layout `state_dict`, so that it can run where
/root/reference is absent (the GPU box).  Each function cites the reference lines it follows; the file is pinned
against fixtures produced by the reference's OWN code (tests/golden/*.pt, made by oracle/make_golden.py) in
tests/test_oracle.py.

All arithmetic in the dtype of the inputs (use fp32 tensors for the oracle role), flips and cats done literally.
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from mamba_ssm.ops.selective_scan_interface import mamba_inner_ref  # noqa: E402
from mamba_ssm.ops.triton.layernorm import _add_norm  # noqa: E402


def mamba_ref(h, sd, prefix):
    """One `Mamba.forward` from state_dict entries `prefix + {in_proj.weight, ...}`."""
    w_in = sd[prefix + "in_proj.weight"]
    xz = torch.einsum("ed,bld->bel", w_in, h)
    b_in = sd.get(prefix + "in_proj.bias")
    if b_in is not None:
        xz = xz + b_in[None, :, None]
    A = -torch.exp(sd[prefix + "A_log"].float())
    return mamba_inner_ref(xz, sd[prefix + "conv1d.weight"], sd.get(prefix + "conv1d.bias"),
                           sd[prefix + "x_proj.weight"], sd[prefix + "dt_proj.weight"], sd[prefix + "out_proj.weight"],
                           sd.get(prefix + "out_proj.bias"), A, None, None, sd[prefix + "D"].float(),
                           delta_bias=sd[prefix + "dt_proj.bias"].float(), delta_softplus=True)


def bimamba_ref(h, sd, prefix, bidirectional=True, strategy="add"):
    """ref:caduceus/modeling_caduceus.py:122-140 — out = M_fwd(h) (+|*) flip(M_rev(flip(h)))."""
    out = mamba_ref(h, sd, prefix + "mamba_fwd.")
    if bidirectional:
        out_rev = mamba_ref(h.flip(1), sd, prefix + "mamba_rev.").flip(1)
        out = out + out_rev if (strategy or "add") == "add" else out * out_rev
    return out


def rc(x):
    """ref:caduceus/modeling_rcps.py:80-83."""
    return torch.flip(x, dims=[-2, -1])


def rcps_wrap(fn, x):
    """ref:caduceus/modeling_rcps.py:85-99."""
    half = x.shape[-1] // 2
    return torch.cat([fn(x[..., :half]), rc(fn(rc(x[..., half:])))], dim=-1)


def norm_ref(x, w, b, residual, eps, is_rms, prenorm, residual_in_fp32):
    return _add_norm(x, w, b, residual, eps, prenorm, residual_in_fp32, is_rms)


def block_ref(h, res, sd, prefix, cfg):
    """One layer.  Non-RCPS: upstream `Block`.  RCPS: ref:caduceus/modeling_rcps.py:160-199,
    including the literal half swap of the fused branch (:177-197)."""
    rcps, fused, is_rms = cfg["rcps"], cfg["fused_add_norm"], cfg["rms_norm"]
    eps, fp32res = cfg["norm_epsilon"], cfg["residual_in_fp32"]
    nprefix = prefix + ("norm." if (fused or not rcps) else "norm.submodule.")
    w, b = sd[nprefix + "weight"], sd.get(nprefix + "bias")
    mix = prefix + ("mixer.submodule." if rcps else "mixer.")
    bim = lambda t: bimamba_ref(t, sd, mix, cfg["bidirectional"], cfg["bidirectional_strategy"])  # noqa: E731

    def plain_norm(v):
        return norm_ref(v, w, b, None, eps, is_rms, False, False)

    if not rcps:
        if not fused:
            res = (h + res) if res is not None else h
            h = plain_norm(res.to(w.dtype))
            if fp32res:
                res = res.float()
        else:
            h, res = norm_ref(h, w, b, res, eps, is_rms, True, fp32res)
        return bim(h), res

    D = h.shape[-1] // 2
    if not fused:       # RCPSAddNormWrapper, ref:caduceus/modeling_rcps.py:107-130
        if res is None:
            res = h
            hn = torch.cat([plain_norm(h[..., :D].to(w.dtype)), rc(plain_norm(rc(h[..., D:]).to(w.dtype)))], dim=-1)
        else:
            r_f = h[..., :D] + res[..., :D]
            r_r = rc(h[..., D:]) + rc(res[..., D:])
            hn = torch.cat([plain_norm(r_f.to(w.dtype)), rc(plain_norm(r_r.to(w.dtype)))], dim=-1)
            res = torch.cat([r_f, rc(r_r)], dim=-1)
        if fp32res:
            res = res.float()
    else:               # fused branch: "fwd" reads [..., D:], "rc" reads [..., :D]
        h_f, r_f = norm_ref(h[..., D:], w, b, res[..., D:] if res is not None else None, eps, is_rms, True, fp32res)
        h_r, r_r = norm_ref(rc(h[..., :D]), w, b, rc(res[..., :D]) if res is not None else None, eps, is_rms, True,
                            fp32res)
        hn = torch.cat([h_f, rc(h_r)], dim=-1)
        res = torch.cat([r_f, rc(r_r)], dim=-1)
    return rcps_wrap(bim, hn), res


def embed_ref(ids, sd, cfg, cmap):
    """ref:caduceus/modeling_caduceus.py:159-163, ref:caduceus/modeling_rcps.py:46-67."""
    base = "caduceus.backbone.embeddings.word_embeddings."
    if not cfg["rcps"]:
        return F.embedding(ids, sd[base + "weight"])
    w = sd[base + "embedding.weight"]
    rc_ids = torch.gather(cmap.unsqueeze(0).expand(ids.shape[0], -1), 1, torch.flip(ids, dims=[-1]))
    return torch.cat([F.embedding(ids, w), torch.flip(F.embedding(rc_ids, w), dims=[-2, -1])], dim=-1)


def final_norm_ref(h, res, sd, cfg):
    """ref:caduceus/modeling_caduceus.py:233-275."""
    rcps, fused, is_rms = cfg["rcps"], cfg["fused_add_norm"], cfg["rms_norm"]
    eps, fp32res = cfg["norm_epsilon"], cfg["residual_in_fp32"]
    base = "caduceus.backbone.norm_f." + ("" if (fused or not rcps) else "submodule.")
    w, b = sd[base + "weight"], sd.get(base + "bias")
    if not fused:
        if not rcps:
            res = (h + res) if res is not None else h
            return norm_ref(res.to(w.dtype), w, b, None, eps, is_rms, False, False)
        D = h.shape[-1] // 2
        r_f = h[..., :D] + res[..., :D]
        r_r = rc(h[..., D:]) + rc(res[..., D:])
        pn = lambda v: norm_ref(v.to(w.dtype), w, b, None, eps, is_rms, False, False)  # noqa: E731
        return torch.cat([pn(r_f), rc(pn(r_r))], dim=-1)
    if not rcps:
        return norm_ref(h, w, b, res, eps, is_rms, False, fp32res)
    D = h.shape[-1] // 2
    h_f = norm_ref(h[..., :D], w, b, res[..., :D], eps, is_rms, False, fp32res)
    h_r = norm_ref(rc(h[..., D:]), w, b, rc(res[..., D:]), eps, is_rms, False, fp32res)
    return torch.cat([h_f, rc(h_r)], dim=-1)


def lm_head_ref(h, sd, cfg, cmap):
    """ref:caduceus/modeling_caduceus.py:474-475, ref:caduceus/modeling_rcps.py:233-246."""
    if not cfg["rcps"]:
        return F.linear(h, sd["lm_head.weight"]).float()
    w = sd["lm_head.lm_head.weight"]
    D = h.shape[-1] // 2
    return (F.linear(h[..., :D], w) + F.linear(torch.flip(h[..., D:], dims=[-1]), w[cmap, :])).float()


def padded_cmap(cfg):
    """Vocabulary padding and identity extension of the complement map, ref:caduceus/modeling_caduceus.py:353-357."""
    V = cfg["vocab_size"]
    m = cfg["pad_vocab_size_multiple"]
    if V % m:
        V += m - V % m
    cm = dict(cfg["complement_map"]) if cfg.get("complement_map") else None
    if cm is not None:
        for i in range(len(cm), V):
            cm[i] = i
        cm = torch.tensor(list(cm.values()), dtype=torch.long)
    return V, cm


def model_ref(ids, sd, cfg, return_hidden=False):
    """CaduceusForMaskedLM.forward(input_ids).logits (ref:caduceus/modeling_caduceus.py:216-276,449-475)."""
    _, cmap = padded_cmap(cfg)
    h = embed_ref(ids, sd, cfg, cmap)
    res = None
    for i in range(cfg["n_layer"]):
        h, res = block_ref(h, res, sd, f"caduceus.backbone.layers.{i}.", cfg)
    h = final_norm_ref(h, res, sd, cfg)
    logits = lm_head_ref(h, sd, cfg, cmap)
    return (logits, h) if return_hidden else logits
