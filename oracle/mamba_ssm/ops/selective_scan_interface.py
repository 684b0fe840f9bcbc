"""ORACLE (test infrastructure only) — CPU restatement of upstream mamba-ssm 1.2.0.post1
`mamba_ssm/ops/selective_scan_interface.py`: `selective_scan_ref`, `mamba_inner_ref`.

Upstream source is NOT vendored under /root/reference (pinned at ref:caduceus_env.yml:49); this file
restates the published algorithm. Parity at this boundary is "unpinned"
(see oracle/README.md): it is cross-checked against transformers' independent `MambaMixer.slow_forward`
and closed-form cases in tests/test_oracle.py.

The reference's call sites for this arithmetic: ref:caduceus/modeling_caduceus.py:105-113,128-133
(through `Mamba.forward`).
"""
import torch
import torch.nn.functional as F


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """Sequential selective scan, all arithmetic in fp32.

    u, delta: (b, d, l);  A: (d, n) real;  B, C: (b, n, l) or (b, g, n, l) time-varying, or (d, n) constant;
    D, delta_bias: (d);  z: (b, d, l).  Returns out (b, d, l) in u.dtype [, last_state (b, d, n) fp32].
    """
    in_dtype = u.dtype
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias.float()[..., None]
    if delta_softplus:
        delta = F.softplus(delta)  # threshold 20, beta 1 (torch default) == upstream kernel
    bsz, dim, seqlen = u.shape
    nstate = A.shape[1]
    A = A.float()

    var_B = B.dim() >= 3
    var_C = C.dim() >= 3
    B = B.float()
    C = C.float()
    if var_B and B.dim() == 3:
        B = B[:, None]          # (b, 1, n, l)
    if var_C and C.dim() == 3:
        C = C[:, None]

    # discretisation: a = exp(delta*A), g = delta*B*u, materialised as (b, d, l, n) like upstream
    a = torch.exp(delta[..., None] * A[None, :, None, :])
    if var_B:
        groups = B.shape[1]
        Bx = B.repeat_interleave(dim // groups, dim=1)            # (b, d, n, l)
        g = delta[..., None] * Bx.permute(0, 1, 3, 2) * u[..., None]
    else:
        g = delta[..., None] * B[None, :, None, :] * u[..., None]
    if var_C:
        groups = C.shape[1]
        Cx = C.repeat_interleave(dim // groups, dim=1).permute(0, 1, 3, 2)   # (b, d, l, n)

    s = u.new_zeros(bsz, dim, nstate)
    ys = []
    for t in range(seqlen):
        s = a[:, :, t] * s + g[:, :, t]
        if var_C:
            ys.append((s * Cx[:, :, t]).sum(-1))
        else:
            ys.append((s * C[None]).sum(-1))
    y = torch.stack(ys, dim=2)
    if D is not None:
        y = y + u * D.float()[None, :, None]
    if z is not None:
        y = y * F.silu(z.float())
    y = y.to(in_dtype)
    return (y, s) if return_last_state else y


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """Upstream's CUDA entry point; the oracle has only the reference semantics."""
    return selective_scan_ref(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def causal_conv1d_ref(x, weight, bias=None, activation=None):
    """Depthwise causal conv (upstream causal_conv1d_ref): x (b, d, l), weight (d, k), bias (d)."""
    in_dtype = x.dtype
    k = weight.shape[-1]
    y = F.conv1d(x.float(), weight.float()[:, None, :], None if bias is None else bias.float(),
                 padding=k - 1, groups=weight.shape[0])[..., : x.shape[-1]]
    if activation in ("silu", "swish"):
        y = F.silu(y)
    return y.to(in_dtype)


def mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                    out_proj_bias, A=None, B=None, C=None, D=None, delta_bias=None, delta_softplus=True):
    """conv+SiLU -> x_proj -> dt_proj -> selective scan -> out_proj, with the same
    rounding points as upstream's pipeline (every intermediate cast to the activation dtype)."""
    seqlen = xz.shape[-1]
    rank = delta_proj_weight.shape[1]
    nstate = A.shape[-1]
    x, z = xz.chunk(2, dim=1)
    w = conv1d_weight.reshape(conv1d_weight.shape[0], conv1d_weight.shape[-1])
    x = causal_conv1d_ref(x, w, conv1d_bias, activation="silu")
    x_dbl = F.linear(x.transpose(1, 2).reshape(-1, x.shape[1]), x_proj_weight)        # (b*l, r+2n)
    delta = (delta_proj_weight @ x_dbl[:, :rank].t()).reshape(delta_proj_weight.shape[0], -1, seqlen)
    delta = delta.transpose(0, 1)                                                       # (b, d, l)
    if B is None:
        B = x_dbl[:, rank:rank + nstate].reshape(-1, seqlen, nstate).transpose(1, 2)    # (b, n, l)
    if C is None:
        C = x_dbl[:, -nstate:].reshape(-1, seqlen, nstate).transpose(1, 2)
    y = selective_scan_ref(x, delta, A, B, C, D, z=z, delta_bias=delta_bias, delta_softplus=delta_softplus)
    return F.linear(y.transpose(1, 2), out_proj_weight, out_proj_bias)


def mamba_inner_fn(*args, **kwargs):
    return mamba_inner_ref(*args, **kwargs)
