"""ORACLE (test infrastructure only) — restatement of the embedding-extraction arithmetic of the reference's VEP script
(SURVEY.md §8f row N4), the biggest inference consumer of the hot path at seq_len 131072:

  * `find_variant_idx`      ref:vep_embeddings.py:170-192  (pure-Python loops, as in the reference)
  * `extract_embeddings`    ref:vep_embeddings.py:278-311  (windowed mean of +-768 tokens around the SNP, ref ++ alt)
  * the strand views        ref:vep_embeddings.py:355-366  (rcps: channel halves, RC half flipped in length and channel;
                                                            non-rcps: a second forward on the RC ids, flipped in length)
Nothing here is imported by the product path."""
import torch

WINDOW_SIZE_BP = 1536          # ref:vep_embeddings.py:26


def find_variant_idx(ref_ids, alt_ids, ref_rc_ids, alt_rc_ids):
    idx = len(ref_ids) // 2
    if ref_ids[idx] == alt_ids[idx]:
        idx = -1
        for i, (r, a) in enumerate(zip(ref_ids, alt_ids)):
            if r != a:
                idx = i
    rc_idx = len(ref_rc_ids) // 2 - 1
    if ref_rc_ids[rc_idx] == alt_rc_ids[rc_idx]:
        rc_idx = -1
        for i, (r, a) in enumerate(zip(ref_rc_ids, alt_rc_ids)):
            if r != a:
                rc_idx = i
    return idx, rc_idx


def extract_embeddings(item_ref, item_alt, variant_idx, bp_per_token=1):
    window_size = WINDOW_SIZE_BP // bp_per_token
    start, end = -window_size // 2, window_size // 2 + 1
    expanded = torch.arange(start, end, device=item_ref.device).unsqueeze(0) + variant_idx.unsqueeze(1).to(item_ref.device)
    expanded = torch.clamp(expanded, 0, item_ref.size(1) - 1)
    gather = lambda t: torch.gather(t, 1, expanded.unsqueeze(-1).expand(-1, -1, t.size(2))).mean(dim=1)   # noqa: E731
    return torch.cat([gather(item_ref), gather(item_alt)], dim=-1)


def strand_views_rcps(output):
    c = output.size(-1)
    return output[..., :c // 2], output[..., c // 2:].contiguous().flip(dims=[1, 2])
