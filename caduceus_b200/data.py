"""GPU-side hg38 batch preparation — the step immediately BEFORE the hot path.

At millions of nucleotides per second per GPU the reference's loader (pyfaidx slice -> per-character Python reverse
complement ref:src/dataloaders/utils/rc.py:17-26 -> HF tokenizer per sequence
ref:src/dataloaders/datasets/hg38_dataset.py:178-192 -> MLM masking ref:src/dataloaders/utils/mlm.py:4-32) becomes the
bottleneck; here the raw FASTA bytes go to the GPU once and one integer kernel produces `(data, target)`.
"""
import ctypes as C

import torch

from . import _lib
from .functional import _launched, _ptr, _require_cuda, _stream


def char_table(tokenizer, device):
    """256-entry byte -> id table of a CaduceusTokenizer: its `_tokenize` upper-cases, unknown characters -> [UNK]."""
    vocab = tokenizer.get_vocab()
    unk = vocab["[UNK]"]
    table = [unk] * 256
    for ch, idx in vocab.items():
        if len(ch) == 1:
            table[ord(ch)] = idx
            table[ord(ch.lower())] = idx
    return torch.tensor(table, dtype=torch.int32, device=device)


def draw_mlm(shape, vocab_len, mlm_probability=0.15, device="cuda", generator=None):
    """The four random tensors of ref:src/dataloaders/utils/mlm.py:14-29, in the reference's order."""
    full = lambda p: torch.full(shape, p, device=device)      # noqa: E731
    masked = torch.bernoulli(full(mlm_probability), generator=generator).bool()
    replaced = torch.bernoulli(full(0.8), generator=generator).bool()
    random_sel = torch.bernoulli(full(0.5), generator=generator).bool()
    random_words = torch.randint(vocab_len, shape, dtype=torch.long, device=device, generator=generator)
    return masked, replaced, random_sel, random_words


def hg38_batch(raw, table, *, n_id, pad_id, mask_id, rc_flags=None, draws=None):
    """raw (B, L) uint8 CUDA tensor of FASTA bytes -> ids (B, L), or (data, target) when MLM `draws` are given."""
    _require_cuda(raw, table)
    lib = _lib.load()
    raw = raw.contiguous()
    B, L = raw.shape
    data = torch.empty(B, L, dtype=torch.long, device=raw.device)
    target = torch.empty_like(data) if draws is not None else None
    u8 = lambda t: None if t is None else t.to(torch.uint8).contiguous()      # noqa: E731
    m, r, s, w = (None, None, None, None) if draws is None else draws
    m, r, s, rc = u8(m), u8(r), u8(s), u8(rc_flags)
    w = None if w is None else w.contiguous()
    a = _lib.Hg38BatchArgs(_ptr(raw), _ptr(rc), _ptr(table), _ptr(m), _ptr(r), _ptr(s), _ptr(w), _ptr(data), _ptr(target),
                           B, L, int(n_id), int(pad_id), int(mask_id))
    _lib.check(lib.cad_hg38_batch_fwd(C.byref(a), _stream()), "cad_hg38_batch_fwd")
    _launched()
    return data if draws is None else (data, target)
