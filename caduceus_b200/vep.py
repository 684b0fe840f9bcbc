"""Variant-effect embedding extraction — the step immediately AFTER the hot path for the
reference's biggest inference consumer, `vep_embeddings.py` at seq_len 131072: last-layer hidden states of the
reference / alternate sequences (and of their reverse complements) pooled over a +-768 bp window around the SNP and
concatenated, in the on-disk format of ref:vep_embeddings.py:334-341,395-399
(`concat_avg_ws`, `rc_concat_avg_ws` next to the metadata columns the caller carries through).

Against the reference loop (ref:vep_embeddings.py:344-385) this
  * runs reference and alternate in ONE batched forward (they differ in one token; the kernels take any batch),
  * never materialises the gathered (B, 1537, C) window nor flipped copies of the (B, L, C) outputs: the window of the
    RC view is the mirrored window of the stored tensor, read in place.
PyTorch ops only (a 1537-row mean is not a kernel worth writing); the model forward is the CUDA hot path."""
import torch

WINDOW_SIZE_BP = 1536          # ref:vep_embeddings.py:26


def find_variant_idx(ref_ids, alt_ids):
    """Token index where `ref_ids` and `alt_ids` (B, L) differ; the LAST difference if several, -1 if none — the
    value the reference's per-example Python loop returns (ref:vep_embeddings.py:170-183), computed on the device."""
    diff = ref_ids != alt_ids
    L = ref_ids.shape[1]
    pos = torch.arange(L, device=ref_ids.device).expand_as(diff)
    last = torch.where(diff, pos, torch.full_like(pos, -1)).max(dim=1).values
    mid = L // 2
    return torch.where(diff[:, mid], torch.full_like(last, mid), last)


def _window_mean(hidden, variant_idx, half, lo_half, flip_len=False, flip_ch=False):
    """Mean over the token window [idx - lo_half, idx + half] (indices clamped to [0, L-1], duplicates counted, as the
    reference's gather does) of `hidden` (B, L, C).  flip_len / flip_ch: the window is taken on the view
    hidden.flip(1) / .flip(2) without building it."""
    B, L, C = hidden.shape
    outs = []
    for b in range(B):
        i = int(variant_idx[b])
        lo, hi = i - lo_half, i + half                        # inclusive, in view coordinates
        n_lo, n_hi = max(0, -lo), max(0, hi - (L - 1))       # clamped duplicates of row 0 / row L-1
        a, z = max(lo, 0), min(hi, L - 1)
        if flip_len:                                          # view row r is stored row L-1-r
            rows = hidden[b, L - 1 - z:L - a] if z >= a else hidden[b, :0]
            first, last = hidden[b, L - 1], hidden[b, 0]
        else:
            rows = hidden[b, a:z + 1] if z >= a else hidden[b, :0]
            first, last = hidden[b, 0], hidden[b, L - 1]
        s = rows.sum(dim=0, dtype=torch.float32) + n_lo * first.float() + n_hi * last.float()
        outs.append(s / float(hi - lo + 1))
    out = torch.stack(outs).to(hidden.dtype)
    return out.flip(-1) if flip_ch else out


def extract_embeddings(item_ref, item_alt, variant_idx, bp_per_token=1, flip_len=False, flip_ch=False):
    """`concat_avg_ws` of ref:vep_embeddings.py:278-311: (B, 2C) = [window mean of ref, window mean of alt]."""
    window = WINDOW_SIZE_BP // bp_per_token
    lo_half, half = -(-window // 2), window // 2             # the reference's start = -window // 2 (floor), end = window // 2 + 1
    return torch.cat([_window_mean(item_ref, variant_idx, half, lo_half, flip_len, flip_ch),
                      _window_mean(item_alt, variant_idx, half, lo_half, flip_len, flip_ch)], dim=-1)


@torch.no_grad()
def variant_embeddings(model, ref_input_ids, alt_input_ids, variant_idx=None, rcps=True, ref_rc_input_ids=None,
                       alt_rc_input_ids=None, bp_per_token=1, autocast_dtype=torch.float16):
    """One batch of the reference's dump loop (ref:vep_embeddings.py:344-385).  `model(ids).last_hidden_state` must be
    (B, L, C) — an `AutoModel` / `caduceus.Caduceus` backbone.  Returns {"concat_avg_ws", "rc_concat_avg_ws"} on the
    model's device; the caller appends its metadata columns (chromosome, labels, distance_to_nearest_tss, tissue_embed)."""
    if variant_idx is None:
        variant_idx = find_variant_idx(ref_input_ids, alt_input_ids)
    B = ref_input_ids.shape[0]
    dev = next(model.parameters()).device
    run = lambda ids: model(ids.to(dev)).last_hidden_state      # noqa: E731
    with torch.autocast(device_type=dev.type, dtype=autocast_dtype, enabled=dev.type == "cuda"):
        both = run(torch.cat([ref_input_ids, alt_input_ids], dim=0))          # one launch sequence for ref and alt
        out_ref, out_alt = both[:B], both[B:]
        if rcps:
            c = out_ref.shape[-1] // 2
            fwd = extract_embeddings(out_ref[..., :c], out_alt[..., :c], variant_idx, bp_per_token)
            # RC view = channel half [c:] flipped in length and channel (ref:vep_embeddings.py:355-360), read in place
            rc = extract_embeddings(out_ref[..., c:], out_alt[..., c:], variant_idx, bp_per_token, flip_len=True, flip_ch=True)
        else:
            assert ref_rc_input_ids is not None and alt_rc_input_ids is not None, "non-rcps models need the RC ids"
            fwd = extract_embeddings(out_ref, out_alt, variant_idx, bp_per_token)
            both_rc = run(torch.cat([ref_rc_input_ids, alt_rc_input_ids], dim=0))
            rc = extract_embeddings(both_rc[:B], both_rc[B:], variant_idx, bp_per_token, flip_len=True)   # :364-366
    return {"concat_avg_ws": fwd, "rc_concat_avg_ws": rc}
