"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, the drop-in API surface
(class names, state_dict keys, config fields, tokenizer ids), and loud failure without CUDA."""
import os
import re

import pytest
import torch

from conftest import ROOT, golden
import caduceus
import caduceus_b200
from caduceus_b200 import _lib
from caduceus_b200 import functional as CF


def test_header_symbols_are_exported_and_bound():
    header = open(os.path.join(ROOT, "include", "caduceus_b200.h")).read()
    declared = set(re.findall(r"\b(cad_[a-z0-9_]+)\s*\(", header))
    declared -= {"cad_dtype"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    if not os.path.exists(_lib.LIB_PATH):
        from caduceus_b200.build import build
        build()
    lib = _lib.load()                      # getattr on every symbol happens inside load()
    assert lib.cad_version() == _lib.ABI_VERSION
    assert lib.cad_scan_chunk_len() == 512


def test_struct_sizes_match_header_layout():
    # field counts guard against silent drift between include/caduceus_b200.h and the ctypes mirror
    header = open(os.path.join(ROOT, "include", "caduceus_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    structs = {m.group(2): m.group(1) for m in re.finditer(r"typedef struct \{([^}]*)\}\s*(\w+);", header)}
    for cname, cls in (("cad_embedding_args", _lib.EmbeddingArgs), ("cad_add_norm_args", _lib.AddNormArgs),
                       ("cad_scan_fwd_args", _lib.ScanFwdArgs), ("cad_conv_fwd_args", _lib.ConvFwdArgs),
                       ("cad_add_norm_bwd_args", _lib.AddNormBwdArgs), ("cad_embedding_bwd_args", _lib.EmbeddingBwdArgs),
                       ("cad_scan_bwd_args", _lib.ScanBwdArgs), ("cad_conv_bwd_args", _lib.ConvBwdArgs),
                       ("cad_conv_xproj_args", _lib.ConvXprojArgs), ("cad_scan_fixup_args", _lib.ScanFixupArgs),
                       ("cad_hg38_batch_args", _lib.Hg38BatchArgs), ("cad_head_ce_args", _lib.HeadCeArgs), ("cad_window_mean_args", _lib.WindowMeanArgs),
                       ("cad_peer_ctx", _lib.PeerCtx),
                       ("cad_scan_adjoint_args", _lib.ScanAdjointArgs)):
        names = []
        for decl in structs[cname].split(";"):
            for part in decl.split(","):
                found = re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())
                if found:
                    names.append(found[0])
        assert names == [f[0] for f in cls._fields_], (cname, names, [f[0] for f in cls._fields_])


def test_cpu_tensors_fail_loudly():
    m = caduceus_b200.BiMambaWrapper(32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CF.add_norm(torch.randn(4, 32), torch.ones(32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CF.embedding(torch.zeros(1, 4, dtype=torch.long), torch.randn(16, 32))


@pytest.mark.parametrize("tag", ["ph_config0", "ps_config0", "ps_nonfused", "ph_nonfused_ln", "ph_mul_untied", "ph_unidir"])
def test_state_dict_keys_and_shapes_match_reference(tag):
    fx = golden(f"model_{tag}.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    model = caduceus.CaduceusForMaskedLM(cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == list(fx["state_dict"].keys())
    for k, v in sd.items():
        assert v.shape == fx["state_dict"][k].shape and v.dtype == fx["state_dict"][k].dtype, k
    model.load_state_dict(fx["state_dict"])
    lm, emb = model.lm_head.weight, model.get_input_embeddings().weight
    assert lm.data_ptr() == emb.data_ptr()          # tied (PS always; Ph under the reference's transformers pin)


def test_init_statistics_follow_reference_init():
    cfg = caduceus.CaduceusConfig(d_model=128, n_layer=4, vocab_size=12, ssm_cfg={"d_state": 16},
                                  initializer_cfg={"initializer_range": 0.02, "rescale_prenorm_residual": True,
                                                   "n_residuals_per_layer": 1})
    m = caduceus.CaduceusForMaskedLM(cfg)
    mf = m.caduceus.backbone.layers[0].mixer.mamba_fwd
    assert torch.allclose(mf.A_log[0, :3], torch.log(torch.tensor([1.0, 2.0, 3.0])))
    assert torch.all(mf.D == 1) and mf.A_log._no_weight_decay and mf.D._no_weight_decay
    dt = torch.nn.functional.softplus(mf.dt_proj.bias)
    assert dt.min() >= 1e-4 and dt.max() <= 0.1001 and mf.dt_proj.bias._no_reinit
    assert m.config.vocab_size == 16
    assert abs(mf.out_proj.weight.std().item() - (1 / (3 * 256) ** 0.5) / 2) < 0.01
    mr = m.caduceus.backbone.layers[0].mixer.mamba_rev
    assert mr.in_proj.weight is mf.in_proj.weight and mr.out_proj.weight is mf.out_proj.weight
    assert mr.x_proj.weight is not mf.x_proj.weight


def test_tokenizer_ids_and_complement_map():
    tok = caduceus.CaduceusTokenizer(model_max_length=32)
    assert tok.get_vocab() == {"[CLS]": 0, "[SEP]": 1, "[BOS]": 2, "[MASK]": 3, "[PAD]": 4, "[RESERVED]": 5,
                               "[UNK]": 6, "A": 7, "C": 8, "G": 9, "T": 10, "N": 11}
    assert list(tok.complement_map.values()) == [0, 1, 2, 3, 4, 5, 6, 10, 9, 8, 7, 11]
    assert tok("acgtn", add_special_tokens=False)["input_ids"] == [7, 8, 9, 10, 11]
    assert tok("AC")["input_ids"] == [7, 8, 1]


def test_config_roundtrip_and_auto_registration(tmp_path):
    from transformers import AutoConfig
    cmap = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 10, 8: 9, 9: 8, 10: 7, 11: 11}
    cfg = caduceus.CaduceusConfig(d_model=64, n_layer=2, vocab_size=12, rcps=True, complement_map=cmap)
    cfg.save_pretrained(tmp_path)
    back = AutoConfig.from_pretrained(tmp_path)
    assert isinstance(back, caduceus.CaduceusConfig) and back.model_type == "caduceus"
    assert list(back.complement_map.values()) == list(cmap.values()) and back.rcps and back.d_model == 64
    for f in ("ssm_cfg", "rms_norm", "residual_in_fp32", "fused_add_norm", "pad_vocab_size_multiple", "norm_epsilon",
              "initializer_cfg", "bidirectional", "bidirectional_strategy", "bidirectional_weight_tie"):
        assert hasattr(back, f)


def test_rcps_lm_head_single_gemm_equals_reference_formula():
    from caduceus_b200.modeling_rcps import RCPSLMHead
    torch.manual_seed(0)
    cmap = {i: i for i in range(16)}
    cmap.update({7: 10, 10: 7, 8: 9, 9: 8})
    head = RCPSLMHead(true_dim=24, vocab_size=16, complement_map=cmap)
    x = torch.randn(2, 9, 48)
    w = head.weight
    ref = torch.nn.functional.linear(x[..., :24], w) + torch.nn.functional.linear(
        torch.flip(x[..., 24:], dims=[-1]), w[head.complement_map, :])
    assert torch.allclose(head(x), ref, atol=1e-5)


def test_job_tables():
    seq, pset, rev = CF.job_tables(2, 2, 2, False, "cpu")
    assert seq.tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert pset.tolist() == [0, 1] * 4
    assert rev.tolist() == [0, 1, 1, 0] * 2          # strand 1: mamba_fwd runs right-to-left, mamba_rev left-to-right
    seq_u, _, _ = CF.job_tables(1, 1, 2, True, "cpu")
    assert seq_u.tolist() == [0, 1]


@pytest.mark.parametrize("rcps", [False, True])
def test_hf_save_pretrained_roundtrip(tmp_path, rcps):
    """HF save_pretrained / from_pretrained keeps every tensor and every tie (head<->table, mamba_rev<->mamba_fwd
    projections) under the installed transformers; the bare backbone loads through AutoModel from the same files."""
    from transformers import AutoModel, AutoModelForMaskedLM
    cmap = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5, 6: 6, 7: 10, 8: 9, 9: 8, 10: 7, 11: 11}
    cfg = caduceus.CaduceusConfig(d_model=64, n_layer=2, vocab_size=12, ssm_cfg={"d_state": 16}, rcps=rcps,
                                  complement_map=dict(cmap) if rcps else None)
    m = caduceus.CaduceusForMaskedLM(cfg)
    m.save_pretrained(tmp_path)
    m2 = AutoModelForMaskedLM.from_pretrained(tmp_path)
    sd1, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd1) == list(sd2) and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    mixer = m2.caduceus.backbone.layers[0].mixer
    mixer = mixer.submodule if rcps else mixer
    assert mixer.mamba_rev.in_proj.weight is mixer.mamba_fwd.in_proj.weight
    assert mixer.mamba_rev.out_proj.weight is mixer.mamba_fwd.out_proj.weight
    assert m2.lm_head.weight.data_ptr() == m2.get_input_embeddings().weight.data_ptr()
    backbone = AutoModel.from_pretrained(tmp_path)
    assert isinstance(backbone, caduceus.Caduceus)
    bsd = backbone.state_dict()
    assert all(torch.equal(bsd[k], sd1["caduceus." + k]) for k in bsd)


def test_argument_marshalling_and_validation_reach_the_library_without_a_gpu(monkeypatch):
    """Every forward-scan variant and the conv_xproj call marshal their argument blocks through ctypes, pass the
    library's own validation and fail LOUDLY at the first CUDA call when there is no device (no fallback); bad
    arguments are rejected by the library before that."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from caduceus_b200 import functional as CF
    from scan_boundary_ref import _problem
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    monkeypatch.setattr(CF, "_stream", lambda: None)
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(100, 64, [(0, 0, 0)], torch.bfloat16, 0)
    packed = (conv_w4, conv_b, dt_b, A2, Dk)
    for v in (None, 0, 3):          # 100 tokens: the per-call choice is the time-parallel kernel, too
        with pytest.raises(RuntimeError, match="cad_bimamba_scan_fwd failed .*cuTensorMapEncodeTiled"):
            CF.scan_fwd(xz, delta, bc, packed, tuple(tabs), 100, variant=v)
    with pytest.raises(RuntimeError, match="variant must be"):
        CF.scan_fwd(xz, delta, bc, packed, tuple(tabs), 100, variant=5)
    cstate = torch.zeros(1, 64, 1, 16)
    with pytest.raises(RuntimeError, match="cad_bimamba_scan_bwd failed .*cuTensorMapEncodeTiled"):
        CF.scan_bwd(xz, delta, bc, torch.zeros_like(delta), packed, tuple(tabs), 100, cstate)
    w_x, w_dt = torch.randn(1, 48, 64).bfloat16(), torch.randn(1, 64, 16).bfloat16()
    for want in (False, True):
        with pytest.raises(RuntimeError, match="cad_conv_xproj_fwd failed"):
            CF.conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, tuple(tabs), 100, want_bcT=want)
    with pytest.raises(RuntimeError, match="cad_bc_transpose failed"):          # variant 20: first launch of its pipeline
        CF.scan_fwd(xz, delta, bc, packed, tuple(tabs), 100, variant=20, nseg=1)
    with pytest.raises(RuntimeError, match="inference only"):
        CF.scan_fwd(xz, delta, bc, packed, tuple(tabs), 100, variant=20, h0=torch.zeros(1, 64, 16))
    with pytest.raises(RuntimeError, match="16-bit I/O"):
        CF.scan_fwd(xz.float(), delta.float(), bc, packed, tuple(tabs), 100, variant=20, nseg=1,
                    bcT=torch.zeros(1, 256, 32))


def test_scan_kernel_choice_per_call():
    """choose_scan_variant: lane = channel (20) for 16-bit calls with >= 8 warps per SM at >= 2048-token segments, else 3."""
    from caduceus_b200 import functional as CF
    bf, f32 = torch.bfloat16, torch.float32
    assert CF.choose_scan_variant(bf, 16, 4, 512, 131072) == 20        # Caduceus-PS, one GPU
    assert CF.choose_scan_variant(bf, 16, 2, 512, 131072) == 20        # Caduceus-Ph, one GPU
    assert CF.choose_scan_variant(bf, 16, 4, 512, 65536) == 20         # PS shard of a 2-way split
    assert CF.choose_scan_variant(bf, 16, 4, 512, 16384) == 3          # PS shard of an 8-way split: too few warps without time parallelism
    assert CF.choose_scan_variant(bf, 16, 4, 512, 1024) == 3
    assert CF.choose_scan_variant(bf, 16, 64, 512, 8192) == 20         # a large batch of medium sequences (1024 warps, no time split)
    assert CF.choose_scan_variant(f32, 16, 4, 512, 131072) == 3
    old = CF.SCAN_VARIANT
    try:
        CF.SCAN_VARIANT = 3
        assert CF.choose_scan_variant(bf, 16, 4, 512, 131072) == 3
        CF.SCAN_VARIANT = 20
        assert CF.choose_scan_variant(bf, 16, 4, 512, 1024) == 20 and CF.choose_scan_variant(f32, 16, 4, 512, 1024) == 3
    finally:
        CF.SCAN_VARIANT = old


def test_segment_count_of_the_lane_per_channel_scan():
    """variant 20: about eight warps per SM (148 SMs when no device is visible), whole 256-token chunks, no segment shorter
    than 2048 tokens; CAD_SCAN_NSEG overrides."""
    from caduceus_b200 import functional as CF
    assert CF.default_nseg(4, 512, 131072, 8) == 18          # Caduceus-PS headline: 4 jobs x 16 channel groups x 18 = 1152 warps
    assert CF.default_nseg(2, 512, 131072, 4) == 37          # Caduceus-Ph: 2 jobs x 16 groups x 37 = 1184 warps
    assert CF.default_nseg(4, 512, 4096, 8) == 2 and CF.default_nseg(4, 512, 1024, 8) == 1
    assert CF.default_nseg(64, 512, 131072, 8) == 1          # a large batch needs no time split
    old = CF.SCAN_NSEG
    try:
        CF.SCAN_NSEG = 9
        assert CF.default_nseg(4, 512, 131072, 8) == 9
    finally:
        CF.SCAN_NSEG = old


def test_pack_w_x_is_the_per_slab_tensor_core_operand():
    """functional.pack_w_x: (P, R+2N, E) -> (P, E/32, 4, 48, 8) = [slab][8-channel group][operand row][channel in group], dt rows
    padded to 16 with zeros, then the B / C rows — the layout cad_conv_xproj_args.w_x_packed documents (include/caduceus_b200.h)
    and csrc/xproj.cu copies slab by slab (3072 bytes each) straight into the K-major UMMA operand."""
    from caduceus_b200 import functional as CF
    for R in (4, 8, 16):
        P, E = 2, 128
        w = torch.arange(P * (R + 32) * E, dtype=torch.float32).view(P, R + 32, E) + 1
        packed = CF.pack_w_x(w.to(torch.bfloat16), R)
        assert packed.shape == (P, E // 32, 4, 48, 8) and packed.is_contiguous()
        assert packed[0, 0].numel() * packed.element_size() == 3072
        wb = w.to(torch.bfloat16)
        for p in range(P):
            for row in range(48):
                src = row if row < R else (None if row < 16 else row - 16 + R)
                got = packed[p, :, :, row, :].reshape(E)          # channel = slab * 32 + group * 8 + j
                want = torch.zeros(E, dtype=torch.bfloat16) if src is None else wb[p, src]
                assert torch.equal(got, want), (R, p, row)


def test_library_sass_holds_the_blackwell_tensor_core_and_tma_paths():
    """The built .so carries what DESIGN.md §4.2 claims for the projection kernel: tcgen05.mma (UTCHMMA), TMEM loads (LDTM), TMA
    tensor loads and stores (UTMALDG / UTMASTG), bulk copies (UBLKCP) — and no legacy warp-level HMMA left in the library."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    from caduceus_b200.build import LIB, build
    build()
    sass = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass


def test_operator_level_shim_binds_the_references_own_model_code(monkeypatch):
    """INTEGRATION.md §2 executed: the reference's OWN caduceus/*.py (loaded from /root/reference, build container only) on top
    of a `mamba_ssm` shim that re-exports caduceus_b200.modules — the five symbols the reference imports
    (ref:caduceus/modeling_caduceus.py:11-27, ref:caduceus/modeling_rcps.py:12-18).  The model constructs with the reference's
    constructor calls, loads a reference-generated checkpoint strictly, is built from this library's operators, and — there being
    no GPU here and no CPU fallback — its forward fails loudly at the first kernel call instead of computing something else."""
    import sys
    import types
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("the reference tree exists only in the build container")
    from caduceus_b200 import modules as M
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import golden
    for name in ("mamba_ssm", "mamba_ssm.modules", "mamba_ssm.ops", "mamba_ssm.ops.triton"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        monkeypatch.setitem(sys.modules, name, pkg)
    simple = types.ModuleType("mamba_ssm.modules.mamba_simple")
    simple.Mamba, simple.Block = M.Mamba, M.Block
    lnorm = types.ModuleType("mamba_ssm.ops.triton.layernorm")
    lnorm.RMSNorm, lnorm.rms_norm_fn, lnorm.layer_norm_fn = M.RMSNorm, M.rms_norm_fn, M.layer_norm_fn
    monkeypatch.setitem(sys.modules, simple.__name__, simple)
    monkeypatch.setitem(sys.modules, lnorm.__name__, lnorm)
    name = "ref_caduceus_on_b200_operators"
    ref = load_reference(name)
    try:
        for tag in ("ps_small", "ph_config0"):       # RCPS: the reference's own RCPSMambaBlock around our Mamba; Ph: our Block too
            fx = golden(f"model_{tag}.pt")
            cfg = ref.configuration_caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
            model = ref.modeling_caduceus.CaduceusForMaskedLM(cfg).eval()
            model.load_state_dict(fx["state_dict"], strict=True)
            assert type(model).__module__ == f"{name}.modeling_caduceus"      # the reference's class, not this repo's drop-in
            assert len(model.caduceus.backbone.layers) == cfg.n_layer
            mambas = [m for m in model.modules() if isinstance(m, M.Mamba)]
            assert len(mambas) >= cfg.n_layer                                 # every mixer is this library's operator
            if not cfg.rcps:
                assert all(isinstance(b, M.Block) for b in model.caduceus.backbone.layers)
            if not torch.cuda.is_available():
                with pytest.raises(RuntimeError, match="no CPU fallback"):
                    with torch.no_grad():
                        model(fx["input_ids"])
    finally:
        for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            del sys.modules[k]
