"""ORACLE (test infrastructure only) — generate tests/golden/*.pt.

Runs the reference's OWN model code (ref:caduceus/modeling_caduceus.py, ref:caduceus/modeling_rcps.py,
loaded verbatim from /root/reference by oracle/ref_loader.py) on top of the CPU `mamba_ssm` restatement in
this directory, in fp32, and stores (config, state_dict, inputs, outputs). Run in the BUILD container
(the GPU box has no /root/reference):

    python oracle/make_golden.py

Every fixture is deterministic (torch.manual_seed) and small enough for git. The weights are perturbed away
from the init (A_log, D, conv bias, norm weights randomised) so that direction-specific and
channel-flip-specific parameters are actually exercised.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ONLY_D118 = "--only-d118" in sys.argv        # add the d_model = 118 fixtures without rewriting the others

SSM_CFG = dict(d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1, dt_init="random",
               dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False, use_fast_path=True)
INIT_CFG = dict(initializer_range=0.02, rescale_prenorm_residual=True, n_residuals_per_layer=1)
# tokenizer complement map, ref:caduceus/tokenization_caduceus.py:49-66
CMAP = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 10, 8: 9, 9: 8, 10: 7, 11: 11}


def hg38_ids(bsz, seqlen, gen):
    """hg38-shaped ids: A,C,G,T uniform (7..10), ~0.5% PAD(4), a few MASK(3)."""
    ids = torch.randint(7, 11, (bsz, seqlen), generator=gen)
    r = torch.rand(bsz, seqlen, generator=gen)
    ids[r < 0.005] = 4
    ids[(r >= 0.005) & (r < 0.125)] = 3
    return ids


def perturb(model, gen):
    """Move the weights off their structured init so tied/untied and flip paths are distinguishable."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("A_log"):
                p.add_(0.3 * torch.randn(p.shape, generator=gen))
            elif name.endswith(".D"):
                p.copy_(1.0 + 0.5 * torch.randn(p.shape, generator=gen))
            elif name.endswith("conv1d.bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=gen))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))
            elif name.endswith("out_proj.weight"):
                p.mul_(4.0)


def make_model_fixture(ref, tag, *, d_model, n_layer, seqlen, bsz, rcps, fused_add_norm, rms_norm=True,
                       residual_in_fp32=False, strategy="add", tie=True, bidirectional=True, seed=0):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    cfg = ref.CaduceusConfig(
        d_model=d_model, n_layer=n_layer, vocab_size=12, ssm_cfg=dict(SSM_CFG), rms_norm=rms_norm,
        fused_add_norm=fused_add_norm, residual_in_fp32=residual_in_fp32, pad_vocab_size_multiple=8,
        norm_epsilon=1e-5, initializer_cfg=dict(INIT_CFG), bidirectional=bidirectional,
        bidirectional_strategy=strategy, bidirectional_weight_tie=tie, rcps=rcps,
        complement_map=dict(CMAP) if rcps else None)
    model = ref.CaduceusForMaskedLM(cfg).eval()
    perturb(model, gen)
    ids = hg38_ids(bsz, seqlen, gen)
    with torch.no_grad():
        out = model(ids, output_hidden_states=True, return_dict=True)
        # NB the reference appends the final normed state to hidden_states only in its fused branch
        # (ref:caduceus/modeling_caduceus.py:274-275), so take last_hidden_state from the backbone itself
        last = model.caduceus(ids, return_dict=True).last_hidden_state
    fx = {
        "config": dict(d_model=d_model, n_layer=n_layer, vocab_size=12, ssm_cfg=dict(SSM_CFG), rms_norm=rms_norm,
                       fused_add_norm=fused_add_norm, residual_in_fp32=residual_in_fp32,
                       pad_vocab_size_multiple=8, norm_epsilon=1e-5, initializer_cfg=dict(INIT_CFG),
                       bidirectional=bidirectional, bidirectional_strategy=strategy,
                       bidirectional_weight_tie=tie, rcps=rcps, complement_map=dict(CMAP) if rcps else None),
        "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
        "input_ids": ids,
        "logits": out.logits.clone(),
        "last_hidden_state": last.clone(),
        "hidden_after_embedding": out.hidden_states[0].clone(),
    }
    path = os.path.join(OUT, f"model_{tag}.pt")
    torch.save(fx, path)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB, logits {tuple(out.logits.shape)}")


def make_mixer_fixture(ref, mc, tag, *, d_model, seqlen, bsz, strategy, tie, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    mixer = mc.BiMambaWrapper(d_model, bidirectional=True, bidirectional_strategy=strategy,
                              bidirectional_weight_tie=tie, **SSM_CFG).eval()
    perturb(mixer, gen)
    h = torch.randn(bsz, seqlen, d_model, generator=gen)
    with torch.no_grad():
        y = mixer(h)
    fx = {"d_model": d_model, "strategy": strategy, "tie": tie, "ssm_cfg": dict(SSM_CFG),
          "state_dict": {k: v.clone() for k, v in mixer.state_dict().items()}, "hidden": h, "out": y}
    path = os.path.join(OUT, f"mixer_{tag}.pt")
    torch.save(fx, path)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB")


def make_block_fixture(ref, mc, tag, *, d_model, seqlen, bsz, rcps, fused, with_residual, seed):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    block = mc.create_block(d_model, ssm_cfg=dict(SSM_CFG), norm_epsilon=1e-5, rms_norm=True,
                            residual_in_fp32=False, fused_add_norm=fused, layer_idx=0, bidirectional=True,
                            bidirectional_strategy="add", bidirectional_weight_tie=True, rcps=rcps).eval()
    perturb(block, gen)
    width = 2 * d_model if rcps else d_model
    h = torch.randn(bsz, seqlen, width, generator=gen)
    res = torch.randn(bsz, seqlen, width, generator=gen) if with_residual else None
    with torch.no_grad():
        y, r = block(h, res)
    fx = {"d_model": d_model, "rcps": rcps, "fused_add_norm": fused, "ssm_cfg": dict(SSM_CFG),
          "state_dict": {k: v.clone() for k, v in block.state_dict().items()},
          "hidden": h, "residual": res, "out_hidden": y, "out_residual": r}
    path = os.path.join(OUT, f"block_{tag}.pt")
    torch.save(fx, path)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB")


def make_scan_fixture(tag, *, bsz, dim, seqlen, nstate, seed, init_ranges):
    from mamba_ssm.ops.selective_scan_interface import selective_scan_ref
    gen = torch.Generator().manual_seed(seed)
    u = torch.randn(bsz, dim, seqlen, generator=gen)
    z = torch.randn(bsz, dim, seqlen, generator=gen)
    Bm = torch.randn(bsz, nstate, seqlen, generator=gen)
    Cm = torch.randn(bsz, nstate, seqlen, generator=gen)
    D = torch.randn(dim, generator=gen)
    if init_ranges:   # init-time ranges: delta in [1e-3, 0.1], A in [-16, -1]
        delta = torch.randn(bsz, dim, seqlen, generator=gen) * 0.5
        dbias = torch.log(torch.expm1(torch.exp(torch.rand(dim, generator=gen) * 4.6 - 6.9)))
        A = -torch.arange(1, nstate + 1, dtype=torch.float32).repeat(dim, 1)
    else:             # upstream test distributions
        delta = 0.5 * torch.rand(bsz, dim, seqlen, generator=gen)
        dbias = 0.5 * torch.rand(dim, generator=gen)
        A = -0.5 * torch.rand(dim, nstate, generator=gen)
    out, last = selective_scan_ref(u, delta, A, Bm, Cm, D, z, dbias, delta_softplus=True, return_last_state=True)
    fx = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=D, z=z, delta_bias=dbias, out=out, last_state=last)
    path = os.path.join(OUT, f"scan_{tag}.pt")
    torch.save(fx, path)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    mc = sys.modules["ref_caduceus.modeling_caduceus"]

    # d_model = 118: the reference's published 1k-token Caduceus-Ph / -PS models (6 slurm scripts of the reference use it) — not a
    # multiple of the 8-element vector width, so the embedding and add+norm kernels take their element-wise instantiation
    make_model_fixture(ref, "ps_d118", d_model=118, n_layer=2, seqlen=300, bsz=1, rcps=True, fused_add_norm=True, seed=9)
    make_model_fixture(ref, "ph_d118", d_model=118, n_layer=2, seqlen=300, bsz=2, rcps=False, fused_add_norm=True, seed=13)
    if ONLY_D118:
        return

    # BASELINE.json configs[0]: D=128, n_layer=4, L=1024, B=1 MLM forward on the CPU reference path
    make_model_fixture(ref, "ph_config0", d_model=128, n_layer=4, seqlen=1024, bsz=1, rcps=False,
                       fused_add_norm=True, seed=0)
    make_model_fixture(ref, "ps_config0", d_model=128, n_layer=4, seqlen=1024, bsz=1, rcps=True,
                       fused_add_norm=True, seed=1)
    # small variants: ragged length (not a multiple of any chunk), batch>1, non-fused norm, LayerNorm,
    # fp32 residual, untied / multiply strategy, unidirectional
    make_model_fixture(ref, "ph_small", d_model=64, n_layer=2, seqlen=333, bsz=2, rcps=False,
                       fused_add_norm=True, seed=2)
    make_model_fixture(ref, "ps_small", d_model=64, n_layer=2, seqlen=333, bsz=2, rcps=True,
                       fused_add_norm=True, seed=3)
    make_model_fixture(ref, "ps_nonfused", d_model=64, n_layer=2, seqlen=257, bsz=2, rcps=True,
                       fused_add_norm=False, seed=4)
    make_model_fixture(ref, "ph_nonfused_ln", d_model=64, n_layer=2, seqlen=257, bsz=2, rcps=False,
                       fused_add_norm=False, rms_norm=False, seed=5)
    make_model_fixture(ref, "ps_ln_fp32res", d_model=64, n_layer=2, seqlen=130, bsz=1, rcps=True,
                       fused_add_norm=True, rms_norm=False, residual_in_fp32=True, seed=6)
    make_model_fixture(ref, "ph_mul_untied", d_model=64, n_layer=2, seqlen=200, bsz=1, rcps=False,
                       fused_add_norm=True, strategy="ew_multiply", tie=False, seed=7)
    make_model_fixture(ref, "ph_unidir", d_model=64, n_layer=2, seqlen=200, bsz=1, rcps=False,
                       fused_add_norm=True, bidirectional=False, strategy=None, seed=8)

    make_mixer_fixture(ref, mc, "add_tied", d_model=64, seqlen=1100, bsz=2, strategy="add", tie=True, seed=10)
    make_mixer_fixture(ref, mc, "mul_untied", d_model=64, seqlen=130, bsz=1, strategy="ew_multiply", tie=False,
                       seed=11)
    make_mixer_fixture(ref, mc, "len1", d_model=64, seqlen=1, bsz=2, strategy="add", tie=True, seed=12)

    for rcps in (False, True):
        for fused in (False, True):
            for with_res in (False, True):
                tag = f"{'ps' if rcps else 'ph'}_{'fused' if fused else 'plain'}_{'res' if with_res else 'nores'}"
                make_block_fixture(ref, mc, tag, d_model=64, seqlen=96, bsz=2, rcps=rcps, fused=fused,
                                   with_residual=with_res, seed=20)

    make_scan_fixture("upstream_ranges", bsz=2, dim=32, seqlen=700, nstate=16, seed=30, init_ranges=False)
    make_scan_fixture("init_ranges", bsz=1, dim=32, seqlen=2500, nstate=16, seed=31, init_ranges=True)


if __name__ == "__main__":
    main()
