"""hg38 batch preparation: oracle vs reference-generated fixture on CPU, CUDA kernel vs both on GPU."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import batch_ref  # noqa: E402


def test_oracle_matches_reference_fixture():
    g = golden("hg38_batch.pt")
    table = batch_ref.char_table(g["vocab"])
    ids = batch_ref.hg38_ids(g["raw"].numpy(), g["rc_flags"].numpy(), table, g["n_id"], g["pad_id"])
    assert torch.equal(ids, g["ids"])
    data, target = batch_ref.mlm_apply(ids, g["draws"], g["pad_id"], g["mask_id"])
    assert torch.equal(data, g["data"]) and torch.equal(target, g["target"])


def test_draw_order_reproduces_reference_rng_stream():
    g = golden("hg38_batch.pt")
    for b, seed in enumerate(g["seeds"]):
        torch.manual_seed(seed)
        d = batch_ref.draw_mlm(g["ids"][b].shape, g["vocab_len"])
        for mine, ref in zip(d, g["draws"]):
            assert torch.equal(mine, ref[b])


def test_package_tokenizer_table_matches_reference_vocab():
    from caduceus_b200.data import char_table
    from caduceus_b200.tokenization_caduceus import CaduceusTokenizer
    g = golden("hg38_batch.pt")
    tok = CaduceusTokenizer(model_max_length=1000)
    assert dict(tok.get_vocab()) == g["vocab"] and len(tok) == g["vocab_len"]
    assert np.array_equal(char_table(tok, "cpu").numpy(), batch_ref.char_table(g["vocab"]))
    assert tok.pad_token_id == g["pad_id"] and tok.convert_tokens_to_ids(tok.mask_token) == g["mask_id"]


def test_reverse_complement_is_an_involution_on_bytes():
    raw = np.random.default_rng(0).integers(0, 256, size=(3, 257), dtype=np.uint8)
    assert np.array_equal(batch_ref.reverse_complement_bytes(batch_ref.reverse_complement_bytes(raw)), raw)


# ---------------------------------------------------------------------------------------------------------- GPU
def _gpu_inputs(g):
    from caduceus_b200.data import char_table
    from caduceus_b200.tokenization_caduceus import CaduceusTokenizer
    tok = CaduceusTokenizer(model_max_length=1000)
    return tok, char_table(tok, "cuda"), dict(n_id=g["n_id"], pad_id=g["pad_id"], mask_id=g["mask_id"])


@pytest.mark.gpu
def test_kernel_matches_reference_fixture_bit_exact():
    from caduceus_b200.data import hg38_batch
    g = golden("hg38_batch.pt")
    _, table, ids_kw = _gpu_inputs(g)
    raw, rc = g["raw"].cuda(), g["rc_flags"].cuda()
    ids = hg38_batch(raw, table, rc_flags=rc, **ids_kw)
    assert torch.equal(ids.cpu(), g["ids"])
    data, target = hg38_batch(raw, table, rc_flags=rc, draws=tuple(d.cuda() for d in g["draws"]), **ids_kw)
    assert torch.equal(data.cpu(), g["data"]) and torch.equal(target.cpu(), g["target"])
    ids_fwd = hg38_batch(raw, table, **ids_kw)                                  # no rc flags at all
    want = batch_ref.hg38_ids(g["raw"].numpy(), np.zeros(len(rc)), batch_ref.char_table(g["vocab"]), g["n_id"], g["pad_id"])
    assert torch.equal(ids_fwd.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("L", [0, 1, 31, 4097])
def test_kernel_ragged_lengths_vs_oracle(L):
    from caduceus_b200.data import hg38_batch
    g = golden("hg38_batch.pt")
    _, table, ids_kw = _gpu_inputs(g)
    rng = np.random.default_rng(L)
    raw = rng.integers(0, 256, size=(3, L), dtype=np.uint8)                     # every byte value, not only letters
    rc = np.array([1, 0, 1], dtype=np.uint8)
    gen = torch.Generator().manual_seed(L)
    draws = batch_ref.draw_mlm((3, L), g["vocab_len"], generator=gen)
    ids = batch_ref.hg38_ids(raw, rc, batch_ref.char_table(g["vocab"]), g["n_id"], g["pad_id"])
    want_d, want_t = batch_ref.mlm_apply(ids, draws, g["pad_id"], g["mask_id"])
    data, target = hg38_batch(torch.from_numpy(raw).cuda(), table, rc_flags=torch.from_numpy(rc).cuda(),
                              draws=tuple(d.cuda() for d in draws), **ids_kw)
    assert torch.equal(data.cpu(), want_d) and torch.equal(target.cpu(), want_t)


@pytest.mark.gpu
def test_kernel_full_length_properties():
    """BASELINE size (L = 131072): RC of the bytes == complement-map + flip of the ids (the identity RCPS relies on,
    ref:caduceus/tokenization_caduceus.py:66-75), and the MLM statistics are the reference's 15 / 80 / 10 / 10."""
    from caduceus_b200.data import draw_mlm, hg38_batch
    g = golden("hg38_batch.pt")
    tok, table, ids_kw = _gpu_inputs(g)
    B, L = 4, 131072
    gen = torch.Generator(device="cuda").manual_seed(3)
    raw = torch.tensor(list(b"ACGTN"), dtype=torch.uint8, device="cuda")[
        torch.randint(0, 5, (B, L), device="cuda", generator=gen)]
    ones = torch.ones(B, dtype=torch.uint8, device="cuda")
    fwd = hg38_batch(raw, table, n_id=-1, pad_id=ids_kw["pad_id"], mask_id=ids_kw["mask_id"])       # keep N as N
    rc = hg38_batch(raw, table, rc_flags=ones, n_id=-1, pad_id=ids_kw["pad_id"], mask_id=ids_kw["mask_id"])
    cmap = torch.tensor([tok.complement_map[i] for i in range(len(tok.complement_map))], device="cuda")
    assert torch.equal(rc, cmap[fwd].flip(-1))
    draws = draw_mlm((B, L), len(tok), generator=gen)
    ids = hg38_batch(raw, table, **ids_kw)
    data, target = hg38_batch(raw, table, draws=draws, **ids_kw)
    m = draws[0]
    assert torch.equal(target[m], ids[m]) and (target[~m] == ids_kw["pad_id"]).all() and torch.equal(data[~m], ids[~m])
    frac_mask = (data[m] == ids_kw["mask_id"]).float().mean().item()
    assert abs(m.float().mean().item() - 0.15) < 0.005 and abs(frac_mask - 0.8) < 0.01
