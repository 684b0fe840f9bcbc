"""ORACLE (test infrastructure only) — CPU restatement of the reference's hg38 batch preparation.

Pinned: `make_golden_batch.py` runs the reference's own functions (string_reverse_complement, CaduceusTokenizer,
mlm_getitem) in the build container and checks this restatement against them bit for bit before writing
tests/golden/hg38_batch.pt.  Only tests/ may import this file.
"""
import numpy as np
import torch

_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in ("AT", "CG", "GC", "TA", "at", "cg", "gc", "ta"):          # ref:src/dataloaders/utils/rc.py:6-14
    _COMP[ord(_a)] = ord(_b)


def reverse_complement_bytes(raw):
    """ref:src/dataloaders/utils/rc.py:17-26 on a uint8 array (last axis = sequence)."""
    return _COMP[np.asarray(raw)[..., ::-1]]


def char_table(vocab):
    """ref:caduceus/tokenization_caduceus.py:90-94: upper-case, then vocab lookup with [UNK] default."""
    table = np.full(256, vocab["[UNK]"], dtype=np.int64)
    for ch, idx in vocab.items():
        if len(ch) == 1:
            table[ord(ch)] = idx
            table[ord(ch.lower())] = idx
    return table


def hg38_ids(raw, rc_flags, table, n_id, pad_id):
    """ref:src/dataloaders/datasets/hg38_dataset.py:172-212 for one batch of equal-length slices."""
    raw = np.asarray(raw).copy()
    for b, flag in enumerate(rc_flags):
        if flag:
            raw[b] = reverse_complement_bytes(raw[b])
    ids = table[raw]
    ids[ids == n_id] = pad_id
    return torch.from_numpy(ids)


def draw_mlm(shape, vocab_len, mlm_probability=0.15, generator=None):
    """The four draws of ref:src/dataloaders/utils/mlm.py:14-29 in the reference's order (CPU generator)."""
    masked = torch.bernoulli(torch.full(shape, mlm_probability), generator=generator).bool()
    replaced = torch.bernoulli(torch.full(shape, 0.8), generator=generator).bool()
    random_sel = torch.bernoulli(torch.full(shape, 0.5), generator=generator).bool()
    words = torch.randint(vocab_len, size=shape, dtype=torch.long, generator=generator)
    return masked, replaced, random_sel, words


def mlm_apply(ids, draws, pad_id, mask_id):
    """ref:src/dataloaders/utils/mlm.py:10-32 with the draws made explicit."""
    masked, replaced, random_sel, words = draws
    data, target = ids.clone(), ids.clone()
    target[~masked] = pad_id
    rep = replaced & masked
    data[rep] = mask_id
    rnd = random_sel & masked & ~rep
    data[rnd] = words[rnd]
    return data, target
