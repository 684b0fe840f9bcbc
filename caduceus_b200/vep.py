"""Variant-effect embedding extraction — the step immediately AFTER the hot path for the
reference's biggest inference consumer, `vep_embeddings.py` at seq_len 131072: last-layer hidden states of the
reference / alternate sequences (and of their reverse complements) pooled over a +-768 bp window around the SNP and
concatenated, in the on-disk format of ref:vep_embeddings.py:334-341,395-399
(`concat_avg_ws`, `rc_concat_avg_ws` next to the metadata columns the caller carries through).

Against the reference loop (ref:vep_embeddings.py:344-385) this
  * runs reference and alternate in ONE batched forward (they differ in one token; the kernels take any batch),
  * never materialises the gathered (B, 1537, C) window nor flipped copies of the (B, L, C) outputs: the window of the
    RC view is the mirrored window of the stored tensor, read in place.
  * pools with one kernel launch per tensor (csrc/vep_pool.cu), variant indices on the device: no host sync per example.
`dump_split` writes the reference's on-disk dictionary.  Not done: restricting the LAST layer's out_proj / norm to the window rows —
the scan of that layer needs the whole sequence anyway, so the saving is ~0.5 % of a forward."""
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
    reference's gather does) of `hidden` (B, L, C) — possibly a channel slice of a wider tensor.  flip_len / flip_ch: the
    window is taken on the view hidden.flip(1) / .flip(2) without building it.
    CUDA tensors: ONE launch of csrc/vep_pool.cu per pooled tensor (variant indices stay on the device: no host sync, no
    per-example loop).  Host tensors (the CPU plumbing tests, embeddings already moved off the GPU): the same window through a
    clamped index + gather in torch."""
    B, L, C = hidden.shape
    variant_idx = variant_idx.to(hidden.device).long()
    if hidden.is_cuda:
        import ctypes as C_
        from . import _lib
        from . import functional as CF
        if hidden.stride(2) != 1 or hidden.stride(0) != L * hidden.stride(1):
            hidden = hidden.contiguous()
        out = torch.empty(B, C, device=hidden.device, dtype=hidden.dtype)
        # a channel slice of a wider (B, L, W) tensor is pooled in place: pitch = stride(1), offset folded into the base pointer
        a = _lib.WindowMeanArgs(CF._ptr(hidden), CF._ptr(variant_idx.contiguous()), CF._ptr(out), B, L, C, hidden.stride(1), 0, C,
                                int(lo_half), int(half), int(flip_len), int(flip_ch), CF._dt(hidden))
        _lib.check(_lib.load().cad_window_mean(C_.byref(a), CF._stream()), "cad_window_mean")
        CF._launched()
        return out
    rows = torch.arange(-lo_half, half + 1, device=hidden.device)[None, :] + variant_idx[:, None]
    rows = rows.clamp(0, L - 1)
    if flip_len:
        rows = L - 1 - rows
    win = torch.gather(hidden, 1, rows[:, :, None].expand(-1, -1, C))
    out = win.mean(dim=1, dtype=torch.float32).to(hidden.dtype)
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


STORAGE_KEYS = ("concat_avg_ws", "rc_concat_avg_ws", "chromosome", "labels", "distance_to_nearest_tss", "tissue_embed")


@torch.no_grad()
def dump_split(model, batches, path=None, rcps=True, bp_per_token=1, autocast_dtype=torch.float16):
    """The reference's per-split dump (ref:vep_embeddings.py:329-399): iterate `batches` (dicts with ref_input_ids, alt_input_ids,
    variant_idx, chromosome, labels, distance_to_nearest_tss, tissue_embed [, ref_rc_input_ids, alt_rc_input_ids]), collect the
    pooled embeddings next to the metadata columns and return — and, with `path`, torch.save — the dictionary
    {concat_avg_ws, rc_concat_avg_ws, chromosome, labels, distance_to_nearest_tss, tissue_embed} of concatenated CPU tensors,
    i.e. the `{split}_embeds_{rank}.pt` file the reference's downstream SVM reads.  Device-to-host copies are non-blocking into
    pinned buffers; the only synchronisation is one at the end."""
    store = {k: [] for k in STORAGE_KEYS}
    for batch in batches:
        for key in ("chromosome", "labels", "distance_to_nearest_tss", "tissue_embed"):
            store[key].append(torch.as_tensor(batch[key]).cpu())
        out = variant_embeddings(model, batch["ref_input_ids"], batch["alt_input_ids"], variant_idx=batch.get("variant_idx"),
                                 rcps=rcps, ref_rc_input_ids=batch.get("ref_rc_input_ids"),
                                 alt_rc_input_ids=batch.get("alt_rc_input_ids"), bp_per_token=bp_per_token,
                                 autocast_dtype=autocast_dtype)
        for key, val in out.items():
            if val.is_cuda:
                host = torch.empty(val.shape, dtype=val.dtype, pin_memory=True)
                host.copy_(val, non_blocking=True)
                val = host
            store[key].append(val)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    result = {k: torch.cat(v, dim=0) if v else torch.empty(0) for k, v in store.items()}
    if path is not None:
        torch.save(result, path)
    return result
