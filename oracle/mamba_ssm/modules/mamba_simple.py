"""ORACLE (test infrastructure only) — CPU restatement of upstream mamba-ssm 1.2.0.post1
`mamba_ssm/modules/mamba_simple.py`: `Mamba`, `Block` (the two symbols the reference imports at
ref:caduceus/modeling_caduceus.py:11-15 and instantiates at :60-83,105-113).

Upstream is not vendored (pin: ref:caduceus_env.yml:49). Parameter names/shapes, init
and the forward pipeline are restated so that the reference's own model code runs on
top of this file and its `state_dict` keys come out identical to a real mamba_ssm install.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from mamba_ssm.ops.selective_scan_interface import mamba_inner_ref
from mamba_ssm.ops.triton.layernorm import RMSNorm, layer_norm_fn, rms_norm_fn


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None):
        super().__init__()
        fk = {"device": device, "dtype": dtype}
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx

        self.in_proj = nn.Linear(d_model, 2 * self.d_inner, bias=bias, **fk)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, kernel_size=d_conv, groups=self.d_inner,
                                padding=d_conv - 1, bias=conv_bias, **fk)
        self.activation = "silu"
        self.act = nn.SiLU()
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **fk)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk)

        # dt_proj.weight preserves variance; dt_proj.bias is softplus^-1 of a log-uniform dt
        std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError(dt_init)
        dt = torch.exp(torch.rand(self.d_inner, **fk) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.dt_proj.bias._no_reinit = True

        # S4D-real init: A[d, n] = -(n + 1)
        A = torch.arange(1, d_state + 1, dtype=torch.float32, device=device).repeat(self.d_inner, 1)
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        self.D._no_weight_decay = True

        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)

    def forward(self, hidden_states, inference_params=None):
        assert inference_params is None, "oracle covers the training/prefill path only"
        bsz, seqlen, _ = hidden_states.shape
        # xz = W_in h^T, laid out (b, 2E, l) like upstream
        xz = (self.in_proj.weight @ hidden_states.reshape(-1, hidden_states.shape[-1]).t())
        xz = xz.reshape(-1, bsz, seqlen).transpose(0, 1)
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(xz.dtype)[None, :, None]
        A = -torch.exp(self.A_log.float())
        return mamba_inner_ref(
            xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
            self.out_proj.weight, self.out_proj.bias, A, None, None, self.D.float(),
            delta_bias=self.dt_proj.bias.float(), delta_softplus=True)

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        raise NotImplementedError("oracle covers the training/prefill path only")


class Block(nn.Module):
    """Add -> Norm -> Mixer, returning (hidden, residual)."""

    def __init__(self, dim, mixer_cls, norm_cls=nn.LayerNorm, fused_add_norm=False, residual_in_fp32=False):
        super().__init__()
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.mixer = mixer_cls(dim)
        self.norm = norm_cls(dim)
        if fused_add_norm:
            assert isinstance(self.norm, (nn.LayerNorm, RMSNorm))

    def forward(self, hidden_states, residual=None, inference_params=None):
        if not self.fused_add_norm:
            residual = (hidden_states + residual) if residual is not None else hidden_states
            hidden_states = self.norm(residual.to(dtype=self.norm.weight.dtype))
            if self.residual_in_fp32:
                residual = residual.to(torch.float32)
        else:
            fn = rms_norm_fn if isinstance(self.norm, RMSNorm) else layer_norm_fn
            hidden_states, residual = fn(hidden_states, self.norm.weight, self.norm.bias, residual=residual,
                                         prenorm=True, residual_in_fp32=self.residual_in_fp32,
                                         eps=self.norm.eps)
        hidden_states = self.mixer(hidden_states, inference_params=inference_params)
        return hidden_states, residual

    def allocate_inference_cache(self, batch_size, max_seqlen, dtype=None, **kwargs):
        return self.mixer.allocate_inference_cache(batch_size, max_seqlen, dtype=dtype, **kwargs)
