"""Generate tests/golden/hg38_batch.pt by running the REFERENCE's own dataloader functions (build container only).

  string_reverse_complement   /root/reference/src/dataloaders/utils/rc.py
  CaduceusTokenizer           /root/reference/caduceus/tokenization_caduceus.py
  mlm_getitem                 /root/reference/src/dataloaders/utils/mlm.py
and the per-item steps of HG38Dataset.__getitem__ (/root/reference/src/dataloaders/datasets/hg38_dataset.py:172-226;
the class itself needs pyfaidx, which is not installed, so its four lines are replayed here on synthetic slices).
Also checks oracle/batch_ref.py against those outputs before writing the fixture.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import batch_ref  # noqa: E402
import ref_loader  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = ref_loader.load_reference()
    R = ref_loader.REFERENCE_DIR
    rc_mod = _load(os.path.join(R, "src/dataloaders/utils/rc.py"), "ref_rc")
    mlm_mod = _load(os.path.join(R, "src/dataloaders/utils/mlm.py"), "ref_mlm")
    Tok = sys.modules["ref_caduceus.tokenization_caduceus"].CaduceusTokenizer
    L = 1000
    tok = Tok(model_max_length=L)
    rng = np.random.default_rng(7)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKMSWBDHVrykm", dtype=np.uint8)
    prob = np.array([20] * 4 + [6] * 4 + [3, 2] + [0.3] * 14, dtype=np.float64)
    B = 6
    raw = rng.choice(alphabet, size=(B, L), p=prob / prob.sum()).astype(np.uint8)
    raw[2, 100:400] = ord("N")                                   # an assembly gap
    raw[4] = np.frombuffer(b"ACGT" * (L // 4), dtype=np.uint8)
    rc_flags = np.array([0, 1, 1, 0, 1, 0], dtype=np.uint8)
    seeds = [11, 12, 13, 14, 15, 16]
    n_id = tok.get_vocab()["N"]
    datas, targets, idss, draws = [], [], [], []
    for b in range(B):
        seq = raw[b].tobytes().decode()
        if rc_flags[b]:
            seq = rc_mod.string_reverse_complement(seq)
        ids = tok(seq, add_special_tokens=False, padding="max_length", max_length=L, truncation=True)["input_ids"]
        ids = torch.LongTensor(ids)
        ids[ids == n_id] = tok.pad_token_id                      # hg38_dataset.py:211-212 (replace_value)
        torch.manual_seed(seeds[b])
        data, target = mlm_mod.mlm_getitem(ids, mlm_probability=0.15, contains_eos=False, tokenizer=tok)
        torch.manual_seed(seeds[b])
        d = batch_ref.draw_mlm(ids.shape, len(tok))
        idss.append(ids); datas.append(data); targets.append(target); draws.append(d)
    ids, data, target = torch.stack(idss), torch.stack(datas), torch.stack(targets)
    draws = tuple(torch.stack([d[i] for d in draws]) for i in range(4))
    vocab = dict(tok.get_vocab())
    table = batch_ref.char_table(vocab)
    mask_id = tok.convert_tokens_to_ids(tok.mask_token)
    # pin the restatement
    ids_o = batch_ref.hg38_ids(raw, rc_flags, table, n_id, tok.pad_token_id)
    assert torch.equal(ids_o, ids), "oracle ids differ from the reference"
    d_o, t_o = batch_ref.mlm_apply(ids_o, draws, tok.pad_token_id, mask_id)
    assert torch.equal(d_o, data) and torch.equal(t_o, target), "oracle MLM differs from the reference"
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "hg38_batch.pt")
    torch.save(dict(raw=torch.from_numpy(raw), rc_flags=torch.from_numpy(rc_flags), seeds=seeds, vocab=vocab,
                    vocab_len=len(tok), n_id=n_id, pad_id=tok.pad_token_id, mask_id=mask_id,
                    ids=ids, data=data, target=target, draws=draws), out)
    frac = (target != tok.pad_token_id).float().mean().item()
    print(f"wrote {out}: B={B} L={L}, masked fraction {frac:.3f}, oracle pinned OK")


if __name__ == "__main__":
    main()
