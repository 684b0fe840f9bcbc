"""ORACLE (test infrastructure only) — CPU restatement of upstream mamba-ssm 1.2.0.post1
`mamba_ssm/ops/triton/layernorm.py`: `RMSNorm`, `rms_norm_fn`, `layer_norm_fn`.

Imported by the reference at ref:caduceus/modeling_caduceus.py:22 and ref:caduceus/modeling_rcps.py:13;
called at ref:caduceus/modeling_caduceus.py:244-273 and ref:caduceus/modeling_rcps.py:177-195 with
non-contiguous (channel-sliced / flipped) views.

Semantics per token row: r = x (+ residual) in fp32; residual_out = r cast to fp32 (residual_in_fp32)
or x.dtype; y = norm(r) * w (+ b) computed in fp32, cast to x.dtype.
"""
import torch
from torch import nn


def _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms):
    out_dtype = x.dtype
    r = x.float()
    if residual is not None:
        r = r + residual.float()
    # upstream: residual_out keeps residual.dtype when a residual is given; residual_in_fp32 only matters for the first block
    res_dtype = residual.dtype if residual is not None else (torch.float32 if residual_in_fp32 else out_dtype)
    # upstream's one-pass kernel stores the sum in res_dtype but normalises the UNROUNDED fp32 sum
    residual_out = r.to(res_dtype)
    if is_rms:
        rstd = torch.rsqrt(r.square().mean(dim=-1, keepdim=True) + eps)
        y = r * rstd
    else:
        mu = r.mean(dim=-1, keepdim=True)
        var = (r - mu).square().mean(dim=-1, keepdim=True)
        y = (r - mu) * torch.rsqrt(var + eps)
    y = y * weight.float()
    if bias is not None:
        y = y + bias.float()
    y = y.to(out_dtype)
    return (y, residual_out) if prenorm else y


def rms_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False):
    return _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms=True)


def layer_norm_fn(x, weight, bias, residual=None, eps=1e-6, prenorm=False, residual_in_fp32=False,
                  is_rms_norm=False):
    return _add_norm(x, weight, bias, residual, eps, prenorm, residual_in_fp32, is_rms=is_rms_norm)


class RMSNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        return rms_norm_fn(x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm,
                           residual_in_fp32=residual_in_fp32)
