"""Throughput of the GPU hg38 batch preparation; run from anywhere on a GPU box."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, time
from caduceus_b200.data import char_table, draw_mlm, hg38_batch
from caduceus_b200.tokenization_caduceus import CaduceusTokenizer
tok = CaduceusTokenizer(model_max_length=131072)
table = char_table(tok, "cuda")
B, L = 8, 131072
raw = torch.tensor(list(b"ACGTNacgt"), dtype=torch.uint8, device="cuda")[torch.randint(0, 9, (B, L), device="cuda")]
rc = (torch.rand(B, device="cuda") < 0.5).to(torch.uint8)
kw = dict(n_id=tok.get_vocab()["N"], pad_id=tok.pad_token_id, mask_id=tok.convert_tokens_to_ids(tok.mask_token))
draws = draw_mlm((B, L), len(tok))
for _ in range(3): hg38_batch(raw, table, rc_flags=rc, draws=draws, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): hg38_batch(raw, table, rc_flags=rc, draws=draws, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"hg38_batch kernel+alloc B={B} L={L}: {ms*1e3:.1f} us/call, {B*L/ms/1e6:.2f} G nt/s, {B*L*28/ms/1e6:.1f} GB/s algorithmic")
e0.record()
for _ in range(20):
    d = draw_mlm((B, L), len(tok)); hg38_batch(raw, table, rc_flags=rc, draws=d, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"with the torch RNG draws: {ms*1e3:.1f} us/call, {B*L/ms/1e6:.2f} G nt/s")
# CPU reference-style loop on one sequence for scale (python per-char RC + table lookup)
import numpy as np
s = raw[0].cpu().numpy().tobytes().decode()
t0 = time.time()
comp = {"A": "T", "C": "G", "G": "C", "T": "A", "a": "t", "c": "g", "g": "c", "t": "a"}
r = "".join(comp.get(ch, ch) for ch in reversed(s))
ids = tok(r, add_special_tokens=False)["input_ids"]
print(f"python RC + HF tokenizer, one 131k sequence: {time.time()-t0:.3f} s")
