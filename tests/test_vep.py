"""CPU tests of the VEP embedding extraction against the oracle restatement of the reference's
arithmetic (oracle/vep_ref.py: ref:vep_embeddings.py:170-192, 278-311, 355-366)."""
import pytest
import torch

import vep_ref as R
from caduceus_b200 import vep as V


@pytest.mark.parametrize("L,idx", [(4000, [2000, 10, 3995, 0, 3999]), (900, [450, 0, 899])])   # windows clamped at both ends
@pytest.mark.parametrize("bp_per_token", [1, 5, 6])
def test_window_means_match_reference_gather(L, idx, bp_per_token):
    g = torch.Generator().manual_seed(L)
    B, C = len(idx), 12
    ref, alt = torch.randn(B, L, C, generator=g), torch.randn(B, L, C, generator=g)
    vi = torch.tensor(idx)
    want = R.extract_embeddings(ref, alt, vi, bp_per_token)
    got = V.extract_embeddings(ref, alt, vi, bp_per_token)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    # the RC view read in place == the reference's materialised flip
    want_rc = R.extract_embeddings(ref.flip(dims=[1, 2]), alt.flip(dims=[1, 2]), vi, bp_per_token)
    got_rc = V.extract_embeddings(ref, alt, vi, bp_per_token, flip_len=True, flip_ch=True)
    assert torch.allclose(got_rc, want_rc, rtol=1e-5, atol=1e-5)
    want_l = R.extract_embeddings(ref.flip(dims=[1]), alt.flip(dims=[1]), vi, bp_per_token)
    assert torch.allclose(V.extract_embeddings(ref, alt, vi, bp_per_token, flip_len=True), want_l, rtol=1e-5, atol=1e-5)


def test_find_variant_idx_matches_reference_loop():
    g = torch.Generator().manual_seed(0)
    L = 64
    ref = torch.randint(7, 11, (6, L), generator=g)
    alt = ref.clone()
    alt[0, L // 2] = 3            # SNP at the midpoint (the reference's first guess)
    alt[1, 5] = 3                 # elsewhere
    alt[2, 7] = 3; alt[2, 40] = 3  # noqa: E702  two differences: the loop keeps the LAST one
    alt[4, 0] = 3
    alt[5, L - 1] = 3             # row 3: identical -> -1
    got = V.find_variant_idx(ref, alt)
    want = [R.find_variant_idx(ref[b].tolist(), alt[b].tolist(), ref[b].tolist(), alt[b].tolist())[0] for b in range(6)]
    assert got.tolist() == want == [L // 2, 5, 40, -1, 0, L - 1]


class _FakeBackbone(torch.nn.Module):
    """last_hidden_state = an embedding lookup: enough to check the batching / strand-view plumbing on CPU."""
    def __init__(self, C):
        super().__init__()
        self.emb = torch.nn.Embedding(16, C)

    def forward(self, ids):
        return type("O", (), {"last_hidden_state": self.emb(ids)})()


@pytest.mark.parametrize("rcps", [True, False])
def test_variant_embeddings_plumbing_matches_reference_loop(rcps):
    torch.manual_seed(1)
    B, L, C = 3, 2000, 8
    m = _FakeBackbone(C)
    ref = torch.randint(7, 11, (B, L))
    alt = ref.clone()
    alt[:, L // 2] = (alt[:, L // 2] - 7 + 1) % 4 + 7      # another base at the midpoint
    comp = torch.tensor([0, 1, 2, 3, 4, 5, 6, 10, 9, 8, 7, 11, 12, 13, 14, 15])
    rc = lambda ids: comp[ids.flip(1)]   # noqa: E731
    out = V.variant_embeddings(m, ref, alt, rcps=rcps, ref_rc_input_ids=rc(ref), alt_rc_input_ids=rc(alt))
    vi = torch.full((B,), L // 2)
    o_ref, o_alt = m(ref).last_hidden_state, m(alt).last_hidden_state
    if rcps:
        f_ref, r_ref = R.strand_views_rcps(o_ref)
        f_alt, r_alt = R.strand_views_rcps(o_alt)
    else:
        f_ref, f_alt = o_ref, o_alt
        r_ref, r_alt = m(rc(ref)).last_hidden_state.flip(dims=[1]), m(rc(alt)).last_hidden_state.flip(dims=[1])
    assert torch.allclose(out["concat_avg_ws"], R.extract_embeddings(f_ref, f_alt, vi), rtol=1e-5, atol=1e-6)
    assert torch.allclose(out["rc_concat_avg_ws"], R.extract_embeddings(r_ref, r_alt, vi), rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_variant_embeddings_on_the_cuda_backbone():
    """The real (rcps) backbone on cuda: the batched ref+alt forward and the in-place RC windows against separate
    forwards pooled by the oracle restatement."""
    import caduceus
    from conftest import golden
    fx = golden("model_ps_small.pt")
    cfg = caduceus.CaduceusConfig(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in fx["config"].items()})
    lm = caduceus.CaduceusForMaskedLM(cfg)
    lm.load_state_dict(fx["state_dict"])
    backbone = lm.caduceus.to("cuda").eval()
    torch.manual_seed(0)
    B, L = 2, 3000
    ref = torch.randint(7, 11, (B, L), device="cuda")
    alt = ref.clone()
    alt[:, L // 2] = (alt[:, L // 2] - 7 + 1) % 4 + 7
    out = V.variant_embeddings(backbone, ref, alt, rcps=True)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        o_ref, o_alt = backbone(ref).last_hidden_state, backbone(alt).last_hidden_state
    vi = torch.full((B,), L // 2, device="cuda")
    f_ref, r_ref = R.strand_views_rcps(o_ref)
    f_alt, r_alt = R.strand_views_rcps(o_alt)
    for key, want in (("concat_avg_ws", R.extract_embeddings(f_ref, f_alt, vi)),
                      ("rc_concat_avg_ws", R.extract_embeddings(r_ref, r_alt, vi))):
        assert out[key].shape == (B, 2 * cfg.d_model)
        assert torch.allclose(out[key].float(), want.float(), rtol=3e-3, atol=5e-3), (key, (out[key] - want).abs().max())


def test_dump_split_writes_the_reference_storage_dict(tmp_path):
    """dump_split == the reference loop's storage_dict after concat_storage_dict_values (ref:vep_embeddings.py:329-399): same keys,
    batches concatenated in order, embeddings equal to the per-batch extraction."""
    torch.manual_seed(2)
    B, L, C, nb = 2, 1700, 8, 3
    m = _FakeBackbone(C)
    batches = []
    for i in range(nb):
        ref = torch.randint(7, 11, (B, L))
        alt = ref.clone()
        alt[:, 100 + i] = (alt[:, 100 + i] - 7 + 1) % 4 + 7
        batches.append({"ref_input_ids": ref, "alt_input_ids": alt, "variant_idx": torch.full((B,), 100 + i),
                        "chromosome": torch.full((B,), i), "labels": torch.tensor([0, 1]), "distance_to_nearest_tss": torch.rand(B),
                        "tissue_embed": torch.full((B,), 7 - i)})
    path = tmp_path / "test_embeds_0.pt"
    got = V.dump_split(m, batches, path=str(path), rcps=True)
    assert tuple(got) == V.STORAGE_KEYS and all(v.shape[0] == nb * B for v in got.values())
    loaded = torch.load(str(path))
    for i, b in enumerate(batches):
        one = V.variant_embeddings(m, b["ref_input_ids"], b["alt_input_ids"], variant_idx=b["variant_idx"], rcps=True)
        for key in ("concat_avg_ws", "rc_concat_avg_ws"):
            assert torch.equal(loaded[key][i * B:(i + 1) * B], one[key])
        assert torch.equal(loaded["chromosome"][i * B:(i + 1) * B], b["chromosome"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_window_mean_kernel_matches_reference_gather(dtype):
    """csrc/vep_pool.cu on a channel slice of a wider tensor, windows clamped at both ends, all four (flip_len, flip_ch) views."""
    g = torch.Generator().manual_seed(5)
    B, L, W = 5, 5000, 48
    hid = torch.randn(B, L, W, generator=g).to("cuda").to(dtype)
    vi = torch.tensor([2500, 3, 4996, 0, 4999], device="cuda")
    for sl in (slice(0, 24), slice(24, 48)):
        for fl in (False, True):
            for fc in (False, True):
                view = hid[..., sl]
                want_src = view.flip(dims=[1]) if fl else view
                want_src = want_src.flip(dims=[2]) if fc else want_src
                want = R.extract_embeddings(want_src.float(), want_src.float(), vi)[:, :24]
                got = V._window_mean(view, vi, 768, 768, flip_len=fl, flip_ch=fc)
                assert torch.allclose(got.float(), want, rtol=2e-3 if dtype == torch.float16 else 1e-5, atol=2e-3 if dtype == torch.float16 else 1e-5)
