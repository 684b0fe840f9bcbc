"""Host-side wrappers over the C-ABI kernels: tensors in, tensors out, autograd where the path trains.

PyTorch is plumbing here (device memory, streams, the dense cuBLAS projections); the arithmetic of the hot
path is in csrc/*.cu.  Every function raises RuntimeError on non-CUDA tensors — there is no CPU fallback.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import CAD_BF16, CAD_F16, CAD_F32

_DT = {torch.float32: CAD_F32, torch.float16: CAD_F16, torch.bfloat16: CAD_BF16}

LAUNCHES = 0          # kernels of THIS library enqueued so far (bench.py reports the count of a timed region)

# Tuning knobs (module globals so that bench.py / tests can flip them; the environment only sets the initial value).
# The two forward-scan kernels are described at cad_scan_fwd_args.variant in include/caduceus_b200.h and measured in DESIGN.md §4.
SCAN_VARIANT = int(os.environ.get("CAD_SCAN_VARIANT", "0"))         # 0 = choose per call (choose_scan_variant); 3 / 20 = force
SCAN_NSEG = int(os.environ.get("CAD_SCAN_NSEG", "0"))               # variant 20: time segments per job (0 = default_nseg)
SCAN_EVENTS = None    # when a list: (start, end) CUDA events are recorded around every fused-scan launch
XPROJ_EVENTS = None   # the same around every conv_xproj launch
# conv_xproj: hand the kernel W_x pre-packed per slab (pack_w_x, one bulk copy per slab) instead of letting it gather 16-byte rows
XPROJ_PACK_W = os.environ.get("CAD_XPROJ_PACK_W", "1") == "1"


def _launched(n=1):
    global LAUNCHES
    LAUNCHES += n


def _dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"caduceus_b200: unsupported dtype {t.dtype} (float32 / float16 / bfloat16 only)")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "caduceus_b200: the hot path runs only on CUDA tensors (sm_100a kernels); there is no CPU fallback. "
                "Use the oracle under oracle/ for CPU reference numbers.")


def round_up(x, m):
    return (x + m - 1) // m * m


def _rows(t, width):
    """View t (..., width) as (rows, width) with a uniform row pitch; copies only when it cannot be a view."""
    if t.stride(-1) != 1:
        t = t.contiguous()
    try:
        v = t.view(-1, width)
    except RuntimeError:
        v = t.contiguous().view(-1, width)
    pitch = v.stride(0) if v.shape[0] > 1 else width
    if pitch < width or pitch % 8 != 0 or v.data_ptr() % 16 != 0:
        v = v.contiguous()
        pitch = width
    return v, v.shape[0], pitch


# =====================================================================================================
# embedding
# =====================================================================================================
class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, weight, cmap):
        _require_cuda(ids, weight, cmap)
        lib = _lib.load()
        ids = ids.contiguous().long()
        w = weight.contiguous()
        B, L = ids.shape
        V, D = w.shape
        rcps = cmap is not None
        out = torch.empty(B, L, 2 * D if rcps else D, device=w.device, dtype=w.dtype)
        a = _lib.EmbeddingArgs(_ptr(ids), _ptr(w), _ptr(cmap), _ptr(out), B, L, V, D, int(rcps), _dt(w))
        _lib.check(lib.cad_embedding_fwd(C.byref(a), _stream()), "cad_embedding_fwd")
        _launched()
        ctx.save_for_backward(ids, cmap)
        ctx.shape = (V, D, w.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        ids, cmap = ctx.saved_tensors
        V, D, wdtype = ctx.shape
        lib = _lib.load()
        dout = dout.contiguous()
        B, L = ids.shape
        dw = torch.zeros(V, D, device=dout.device, dtype=torch.float32)
        a = _lib.EmbeddingBwdArgs(_ptr(ids), _ptr(cmap), _ptr(dout), _ptr(dw), B, L, V, D, int(cmap is not None),
                                  _dt(dout))
        _lib.check(lib.cad_embedding_bwd(C.byref(a), _stream()), "cad_embedding_bwd")
        _launched()
        return None, dw.to(wdtype), None


def embedding(ids, weight, cmap=None):
    """Ph: W[ids].  PS (cmap given): cat[W[ids], W[cmap[ids]][..., ::-1]]  ==  ref:caduceus/modeling_rcps.py:54-67."""
    return _EmbeddingFn.apply(ids, weight, cmap)


# =====================================================================================================
# fused add + RMSNorm / LayerNorm
# =====================================================================================================
class _AddNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, is_rms, want_res, residual_in_fp32, nhalf, swap, wflip_mask):
        _require_cuda(x, residual, weight, bias)
        lib = _lib.load()
        width = x.shape[-1]
        D = width // nhalf
        shape = x.shape
        x2, rows, ldx = _rows(x, width)
        r2, ldr = None, 0
        if residual is not None:
            assert residual.shape == x.shape
            r2, _, ldr = _rows(residual, width)
        w = weight.contiguous()
        b = bias.contiguous() if bias is not None else None
        if b is not None and b.dtype != w.dtype:
            b = b.to(w.dtype)
        needs_grad = any(ctx.needs_input_grad[:4])
        # upstream keeps the dtype of a residual that is passed in; residual_in_fp32 only decides the dtype of a NEW residual stream
        res_dtype = residual.dtype if residual is not None else (torch.float32 if residual_in_fp32 else x.dtype)
        y = torch.empty(shape, device=x.device, dtype=x.dtype)
        # the pre-norm sum is written when the caller wants it, or (training) to feed the backward
        store_res = want_res or needs_grad
        res_out = torch.empty(shape, device=x.device, dtype=res_dtype) if store_res else None
        rstd = torch.empty(rows * nhalf, device=x.device, dtype=torch.float32) if needs_grad else None
        mean = torch.empty(rows * nhalf, device=x.device, dtype=torch.float32) if (needs_grad and not is_rms) else None
        a = _lib.AddNormArgs(
            _ptr(x2), _ptr(r2), _ptr(w), _ptr(b), _ptr(y), _ptr(res_out), _ptr(rstd), _ptr(mean),
            rows, D, ldx, ldr, width, width, nhalf, swap, wflip_mask, int(is_rms),
            _dt(x2), _dt(w), _dt(r2) if r2 is not None else CAD_F32, _DT[res_dtype], float(eps))
        _lib.check(lib.cad_add_norm_fwd(C.byref(a), _stream()), "cad_add_norm_fwd")
        _launched()
        if needs_grad:
            ctx.save_for_backward(res_out, w, rstd, mean)
            ctx.meta = (rows, D, width, nhalf, swap, wflip_mask, is_rms, b is not None, x.dtype,
                        None if residual is None else residual.dtype, weight.dtype,
                        None if bias is None else bias.dtype, shape)
        if want_res:
            return y, res_out
        return y, None

    @staticmethod
    def backward(ctx, dy, dres):
        res_out, w, rstd, mean = ctx.saved_tensors
        rows, D, width, nhalf, swap, wflip_mask, is_rms, has_bias, xdtype, rdtype, wdtype, bdtype, shape = ctx.meta
        lib = _lib.load()
        if dy is None:
            dy = torch.zeros(shape, device=res_out.device, dtype=xdtype)
        dy2, _, lddy = _rows(dy, width)
        dr2, lddr = None, 0
        if dres is not None:
            dr2, _, lddr = _rows(dres, width)
        # gradient of the (x + residual) sum; kept in fp32 when the residual stream is fp32
        gdtype = torch.float32 if (rdtype == torch.float32 or res_out.dtype == torch.float32) else xdtype
        dx = torch.empty(shape, device=dy.device, dtype=gdtype)
        nblocks = lib.cad_add_norm_bwd_blocks(rows * nhalf)
        dwp = torch.empty(nblocks, D, device=dy.device, dtype=torch.float32)
        dbp = torch.empty(nblocks, D, device=dy.device, dtype=torch.float32) if has_bias else None
        a = _lib.AddNormBwdArgs(
            _ptr(dy2), _ptr(dr2), _ptr(res_out), _ptr(w), _ptr(rstd), _ptr(mean), _ptr(dx), _ptr(dwp), _ptr(dbp),
            rows, D, lddy, lddr, width, width, nhalf, swap, wflip_mask, int(is_rms), int(has_bias),
            _dt(dy2), _dt(res_out), _dt(w), _DT[gdtype], _dt(dr2) if dr2 is not None else CAD_F32, nblocks)
        _lib.check(lib.cad_add_norm_bwd(C.byref(a), _stream()), "cad_add_norm_bwd")
        _launched()
        dw = dwp.sum(0).to(wdtype)
        db = dbp.sum(0).to(bdtype) if has_bias else None
        gx = dx if dx.dtype == xdtype else dx.to(xdtype)
        gr = None
        if rdtype is not None:
            gr = dx if dx.dtype == rdtype else dx.to(rdtype)
        return gx, gr, dw, db, None, None, None, None, None, None, None


def add_norm(x, weight, bias=None, residual=None, eps=1e-5, is_rms=True, prenorm=False, residual_in_fp32=False,
             nhalf=1, swap=0, wflip_mask=0):
    """y = norm(x + residual) * w (+ b) per D-wide half; returns y or (y, x + residual).

    nhalf=2 covers the RC-equivariant blocks with no flips or cats (see include/caduceus_b200.h):
      fused RCPSMambaBlock (ref:caduceus/modeling_rcps.py:174-197): swap=1, wflip_mask=2
      RCPSAddNormWrapper / final norm (ref:caduceus/modeling_rcps.py:107-130,
      ref:caduceus/modeling_caduceus.py:242-262): swap=0, wflip_mask=2
    """
    y, res = _AddNormFn.apply(x, residual, weight, bias, eps, is_rms, prenorm, residual_in_fp32, nhalf, swap,
                              wflip_mask)
    return (y, res) if prenorm else y


# =====================================================================================================
# BiMamba inner path
# =====================================================================================================
_JOB_CACHE = {}
_JOB_ADJACENT = {}
JOB_HOST = {}


def job_tables(nbatch, nstrand, ndir, untied_in, device):
    """Device int32 tables (seq_of_job, pset_of_job, rev_of_job) for jobs ordered (batch, strand, direction).

    Strand 1 is the reverse-complement strand of RCPSWrapper (ref:caduceus/modeling_rcps.py:95-99): its input is
    time-reversed, so mamba_fwd's parameters run right-to-left there and mamba_rev's left-to-right
    (SURVEY.md A.6): rev = direction XOR strand.
    """
    key = (nbatch, nstrand, ndir, untied_in, str(device))
    hit = _JOB_CACHE.get(key)
    if hit is not None:
        return hit
    seq, pset, rev = [], [], []
    nw = ndir if untied_in else 1
    for b in range(nbatch):
        for s in range(nstrand):
            for d in range(ndir):
                seq.append((b * nstrand + s) * nw + (d if untied_in else 0))
                pset.append(d)
                rev.append(d ^ s)
    t = tuple(torch.tensor(v, dtype=torch.int32, device=device) for v in (seq, pset, rev))
    _JOB_CACHE[key] = t
    # host-side fact used by the backward: are the jobs of one in-proj output adjacent and equally many?
    nseq = nbatch * nstrand * nw
    _JOB_ADJACENT[t[0].data_ptr()] = (len(seq) % nseq == 0 and
                                      seq == [j // (len(seq) // nseq) for j in range(len(seq))])
    JOB_HOST[t[0].data_ptr()] = (tuple(seq), tuple(pset), tuple(rev))      # host copies: no device sync to read them
    return t


def conv_silu(xz, conv_w4, conv_b, jobs, L, halo=None):
    """v1 helper: u[job] = silu((anti)causal conv of x rows) -> (njobs, E, ldxz)."""
    lib = _lib.load()
    seq, pset, rev = jobs
    nseq, twoE, ld = xz.shape
    E = twoE // 2
    njobs = seq.numel()
    u = torch.empty(njobs, E, ld, device=xz.device, dtype=xz.dtype)
    a = _lib.ConvFwdArgs(_ptr(xz), _ptr(u), _ptr(conv_w4), _ptr(conv_b), _ptr(seq), _ptr(pset), _ptr(rev), _ptr(halo),
                         L, E, ld, ld, nseq, njobs, _dt(xz))
    _lib.check(lib.cad_conv_silu_fwd(C.byref(a), _stream()), "cad_conv_silu_fwd")
    _launched()
    return u


def conv_xproj_supported(xz, N, R):
    E = xz.shape[1] // 2
    return xz.dtype in (torch.bfloat16, torch.float16) and N == 16 and 1 <= R <= 16 and E % 64 == 0 and E <= 2048


def pack_w_x(w_x, R):
    """w_x (P, R+2N, E) -> the per-slab K-major tensor-core operand cad_conv_xproj_args.w_x_packed describes, (P, E/32, 4, 48, 8):
    dt rows padded to 16 with zeros, then the B / C rows; done once per weight version (derived-weight cache)."""
    P, rows, E = w_x.shape
    pad = torch.zeros(P, 48, E, device=w_x.device, dtype=w_x.dtype)
    pad[:, :R] = w_x[:, :R]
    pad[:, 16:16 + rows - R] = w_x[:, R:]
    return pad.view(P, 48, E // 32, 4, 8).permute(0, 2, 3, 1, 4).contiguous()


def conv_xproj(xz, w_x, w_dt, conv_w4, conv_b, jobs, L, halo=None, want_bcT=False, w_x_packed=None):
    """Fused conv+SiLU -> x_proj -> dt_proj on tensor cores (tcgen05 + TMEM): returns (delta (njobs, E, ld), bc (njobs, 2N, ldbc)
    fp32) without materialising u.  w_x (P, R+2N, E), w_dt (P, E, R) in the activation dtype.
    want_bcT: also return the B / C rows token-major, (njobs, ceil256(L), 2N) fp32 — what scan variant 20 reads."""
    lib = _lib.load()
    seq, pset, rev = jobs
    nseq, twoE, ld = xz.shape
    E = twoE // 2
    R = w_dt.shape[-1]
    N = (w_x.shape[1] - R) // 2
    njobs = seq.numel()
    ldbc = round_up(max(L, 1), 32)
    delta = torch.empty(njobs, E, ld, device=xz.device, dtype=xz.dtype)
    bc = torch.empty(njobs, 2 * N, ldbc, device=xz.device, dtype=torch.float32)
    a = _lib.ConvXprojArgs(_ptr(xz), _ptr(w_x.contiguous()), _ptr(w_dt.contiguous()), _ptr(conv_w4), _ptr(conv_b),
                           _ptr(seq), _ptr(pset), _ptr(rev), _ptr(halo), _ptr(delta), _ptr(bc),
                           L, E, N, R, ld, ld, ldbc, nseq, njobs, _dt(xz), None, 0, _ptr(w_x_packed))
    bcT = None
    if want_bcT:
        Lp, L128 = round_up(max(L, 1), 256), round_up(max(L, 1), 128)
        bcT = torch.empty(njobs, Lp, 2 * N, device=xz.device, dtype=torch.float32)
        if Lp > L128:
            bcT[:, L128:].zero_()                 # rows no 128-token tile of the kernel covers
        a.bcT, a.ldT = _ptr(bcT), Lp
    ev = None
    if XPROJ_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.check(lib.cad_conv_xproj_fwd(C.byref(a), _stream()), "cad_conv_xproj_fwd")
    _launched()
    if ev is not None:
        ev[1].record()
        XPROJ_EVENTS.append(ev)
    return (delta, bc, bcT) if want_bcT else (delta, bc)


def project_dt_bc(xdbl, dt_w_job, L, N):
    """From x_dbl (njobs, R+2N, ld): dt_raw = W_dt . x_dbl[:R] (a K=R GEMM, io dtype) and the fp32 B/C rows in
    the TMA-friendly pitch (multiple of 32 tokens, zero padded)."""
    njobs, rows, ld = xdbl.shape
    R = rows - 2 * N
    delta = torch.bmm(dt_w_job, xdbl[:, :R, :])                                   # (njobs, E, ld)
    ldbc = round_up(max(L, 1), 32)
    bc = torch.zeros(njobs, 2 * N, ldbc, device=xdbl.device, dtype=torch.float32)
    bc[:, :, :L] = xdbl[:, R:, :L]
    return delta, bc


def scan_fwd(xz, delta, bc, packed, jobs, L, *, halo=None, h0=None, want_state=False, want_chunk_state=False,
             channels_per_cta=0, state_only=False, variant=None, nseg=None, bcT=None):
    """Launch the fused bidirectional scan.
    xz (nseq, 2E, ld), delta (njobs, E, ld), bc (njobs, 2N, ldbc) fp32 -> out (njobs, E, ld).
    variant: None = SCAN_VARIANT, else choose_scan_variant(); nseg: time segments per job for variant 20 (None = default_nseg)."""
    lib = _lib.load()
    seq, pset, rev = jobs
    conv_w4, conv_b, dt_b, A2, Dk = packed
    nseq, twoE, ldxz = xz.shape
    E = twoE // 2
    njobs, twoN, ldbc = bc.shape
    P, _, N = A2.shape
    assert twoN == 2 * N and delta.shape[0] == njobs and delta.shape[1] == E
    want_state = want_state or state_only
    plain = h0 is None and not want_chunk_state and not state_only
    if variant is None:
        variant = choose_scan_variant(xz.dtype, N, njobs, E, L) if plain else 3
    variant = int(variant) or 3
    out = None if state_only else torch.empty(njobs, E, ldxz, device=xz.device, dtype=xz.dtype)
    hlast = torch.empty(njobs, E, N, device=xz.device, dtype=torch.float32) if want_state else None
    dtsum = torch.empty(njobs, E, device=xz.device, dtype=torch.float32) if want_state else None
    chunk = lib.cad_scan_chunk_len()
    nchunks = (L + chunk - 1) // chunk
    cstate = (torch.empty(njobs, E, nchunks, N, device=xz.device, dtype=torch.float32)
              if want_chunk_state else None)
    # a zero-carry shard scan (want_state, no carry-in) on the time-parallel kernel also reports the sum of dt per 512-token chunk,
    # so that scan_fixup can apply the shard's carry-in to several segments in parallel (shard_fixup_plan)
    plan = shard_fixup_plan(L) if (variant == 3 and want_state and h0 is None and not state_only and not want_chunk_state) else None
    chunk_dt = torch.empty(njobs, E, nchunks, device=xz.device, dtype=torch.float32) if plan is not None else None
    a = _lib.ScanFwdArgs(
        _ptr(xz), _ptr(delta), _ptr(bc), _ptr(out), _ptr(conv_w4), _ptr(conv_b), _ptr(dt_b), _ptr(A2), _ptr(Dk),
        _ptr(seq), _ptr(pset), _ptr(rev), _ptr(halo), _ptr(h0), _ptr(hlast), _ptr(dtsum), _ptr(cstate),
        L, E, N, 4, ldxz, delta.stride(1), ldbc, ldxz, nseq, njobs, P, _dt(xz), channels_per_cta, int(state_only), variant)
    a.chunk_dtsum = _ptr(chunk_dt)
    if variant == 20:
        if not plain:
            raise RuntimeError("scan variant 20 covers inference only (no carry-in at launch, saved chunk states or state-only "
                               "pass; a sequence shard takes its carry-in through scan_fixup(..., seg_ctx=...))")
        # the kernel itself produces neither end state nor sum dt: they are composed from the segment outputs
        a.hlast, a.dtsum = None, None
        return scan_fwd_segmented(xz, delta, bc, packed, jobs, L, out, a, nseg=nseg, warps_per_cta=channels_per_cta,
                                  bcT=bcT, want_state=want_state)
    ev = None
    if SCAN_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.check(lib.cad_bimamba_scan_fwd(C.byref(a), _stream()), "cad_bimamba_scan_fwd")
    _launched()
    if ev is not None:
        ev[1].record()
        SCAN_EVENTS.append(ev)
    if plan is not None:
        return out, hlast, dtsum, {"chunk_dtsum": chunk_dt, "nseg": plan[0], "per512": plan[1], "nchunks": nchunks}
    return out, hlast, dtsum, cstate


def shard_fixup_plan(L):
    """(nseg, per512) for the segment-parallel carry fix-up of a shard of L tokens scanned by the time-parallel kernel, or None when
    the serial whole-shard form is used: segments of `per512` physical 512-token chunks (>= 2048 tokens, at most 16 segments) that the
    fix-up kernel's own block geometry (ceil(ceil(L / 256) / nseg) 256-token chunks per block) reproduces exactly."""
    nch512, nch256 = (L + 511) // 512, (L + 255) // 256
    if nch512 < 8:
        return None
    per512 = max(4, -(-nch512 // 16))
    nseg = -(-nch512 // per512)
    if nseg < 2 or -(-nch256 // nseg) != 2 * per512:
        return None
    return nseg, per512


def v20_warps_per_cta(njobs, E):
    """Warps (32 channels each) per CTA of scan variant 20: fewer when there are few jobs, so that the segments fill two CTAs per SM."""
    return min(8 if njobs >= 4 else 4 if njobs >= 2 else 2, (E + 31) // 32)


def choose_scan_variant(dtype, N, njobs, E, L):
    """Forward-scan kernel for a plain inference call (no carry-in / saved states): SCAN_VARIANT when forced, else the lane = channel
    kernel (20) when the call is 16-bit and long enough to give every SM at least six of its warps with segments of >= 2048
    tokens (Caduceus-PS and -Ph at 131k on one GPU, halves of it on two), else the time-parallel kernel (3): short sequences and
    the shards of a 4- or 8-way split have too few (job, channel group, segment) warps for a kernel without time parallelism."""
    ok20 = dtype in (torch.bfloat16, torch.float16) and N == 16
    if SCAN_VARIANT:
        return SCAN_VARIANT if (SCAN_VARIANT != 20 or ok20) else 3
    if not ok20:
        return 3
    sms = _lib.load().cad_sm_count() or 148
    warps = njobs * ((E + 31) // 32) * default_nseg(njobs, E, L, v20_warps_per_cta(njobs, E))
    return 20 if warps >= 6 * sms else 3


def default_nseg(njobs, E, L, warps_per_cta=8):
    """Segments per job for scan variant 20: enough (job, channel group, segment) WARPS for eight per SM — measured optimum on the
    Caduceus-PS / -Ph headline shapes (profiles/r2_call6.log: PS 18 segments 1.87 ms vs 37 segments 1.99 ms; every segment boundary
    costs a carry fix-up, fewer warps cost MUFU occupancy in the scan) — in whole 256-token chunks and never shorter than 2048 tokens."""
    if SCAN_NSEG > 0:
        return SCAN_NSEG
    sms = _lib.load().cad_sm_count() or 148
    want = max(1, (8 * sms) // max(1, njobs * ((E + 31) // 32)))
    return max(1, min(want, L // 2048))


def scan_fwd_segmented(xz, delta, bc, packed, jobs, L, out, a, nseg=None, warps_per_cta=0, cutoff_log2=None, bcT=None,
                       want_state=False):
    """Scan variant 20 (lane = channel, csrc/scan_fwd_v20.cuh): token-major copy of B / C, every segment scanned from a zero
    state, carries composed (cad_seg_carry) and added in place by the segment mode of the fix-up kernel.  `a` is the
    marshalled argument block of scan_fwd (reused so that the two paths cannot drift apart).  cutoff_log2: a carry term is
    dropped once its decay factor is below 2^cutoff (None = default_cutoff_log2(dtype)).
    Returns (out, hlast, dtsum, seg_ctx).  want_state (a sequence SHARD): the local carries are NOT applied
    here; hlast / dtsum are the shard's zero-carry end state and sum dt for the all_gather, and seg_ctx goes to
    scan_fixup(..., h0, seg_ctx=seg_ctx), which applies the shard's carry-in and the local carries in ONE pass."""
    lib = _lib.load()
    seq, pset, rev = jobs
    conv_w4, conv_b, dt_b, A2, Dk = packed
    njobs, twoN, ldbc = bc.shape
    E, N, dev = a.E, twoN // 2, xz.device
    if cutoff_log2 is None:
        cutoff_log2 = default_cutoff_log2(xz.dtype)
    W = warps_per_cta if warps_per_cta > 0 else v20_warps_per_cta(njobs, E)
    nseg = default_nseg(njobs, E, L, W) if nseg is None else int(nseg)
    Lp = round_up(max(L, 1), 256)
    if bcT is None:                               # conv_xproj(want_bcT=True) writes it directly; otherwise one transpose launch
        bcT = torch.empty(njobs, Lp, twoN, device=dev, dtype=torch.float32)
        _lib.check(lib.cad_bc_transpose(_ptr(bc), _ptr(bcT), njobs, twoN, L, ldbc, _stream()), "cad_bc_transpose")
        _launched()
    assert bcT.shape == (njobs, Lp, twoN) and bcT.is_contiguous()
    seg_state = torch.empty(njobs, nseg, E, N, device=dev, dtype=torch.float32)
    seg_dtsum = torch.empty(njobs, nseg, E, device=dev, dtype=torch.float32)
    a.bcT, a.nseg, a.seg_state, a.seg_dtsum, a.channels_per_cta = _ptr(bcT), nseg, _ptr(seg_state), _ptr(seg_dtsum), W
    ev = None
    if SCAN_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.check(lib.cad_bimamba_scan_fwd(C.byref(a), _stream()), "cad_bimamba_scan_fwd")
    _launched()
    hlast = dtsum = None
    if want_state:
        hlast = torch.empty(njobs, E, N, device=dev, dtype=torch.float32)
        dtsum = torch.empty(njobs, E, device=dev, dtype=torch.float32)
        _lib.check(lib.cad_seg_carry(_ptr(seg_state), _ptr(seg_dtsum), _ptr(A2), _ptr(pset), None, None, _ptr(hlast),
                                     _ptr(dtsum), njobs, nseg, E, _stream()), "cad_seg_carry")
        _launched()
    elif nseg > 1:
        carry = torch.empty(njobs, nseg, E, N, device=dev, dtype=torch.float32)
        _lib.check(lib.cad_seg_carry(_ptr(seg_state), _ptr(seg_dtsum), _ptr(A2), _ptr(pset), None, _ptr(carry), None, None,
                                     njobs, nseg, E, _stream()), "cad_seg_carry")
        f = _lib.ScanFixupArgs(_ptr(xz), _ptr(delta), _ptr(bc), _ptr(out), _ptr(dt_b), _ptr(A2), _ptr(seq), _ptr(pset),
                               _ptr(rev), None, L, E, N, a.ldxz, a.ldd, ldbc, a.ldo, a.nseq, njobs, a.io_dtype, 0,
                               float(cutoff_log2), nseg, _ptr(carry))
        _lib.check(lib.cad_bimamba_scan_fixup(C.byref(f), _stream()), "cad_bimamba_scan_fixup")
        _launched(2)
    if ev is not None:
        ev[1].record()
        SCAN_EVENTS.append(ev)
    seg_ctx = {"nseg": nseg, "seg_state": seg_state, "seg_dtsum": seg_dtsum, "cutoff_log2": cutoff_log2} if want_state else None
    return out, hlast, dtsum, seg_ctx


def default_cutoff_log2(dtype):
    """Carry terms are dropped once their decay factor is below 2^cutoff: 2^-16 for bf16 and 2^-20 for fp16 outputs — with all 16
    states of a channel at the threshold and |C h0| as large as the output itself that is 1/16 (1/32) of an output ulp — 2^-40 for fp32."""
    return -16.0 if dtype == torch.bfloat16 else -20.0 if dtype == torch.float16 else -40.0


def scan_fixup(xz, delta, bc, out, packed, jobs, L, h0, cutoff_log2=None, channels_per_cta=0, seg_ctx=None):
    """In place: out += silu(z) * sum_n C * exp2(A2 * cumsum(dt)) * h0 — turns a zero-carry shard scan into the scan
    with carry-in h0 (njobs, E, N).  See csrc/scan_fixup.cu.  cutoff_log2: None = default_cutoff_log2(dtype)."""
    lib = _lib.load()
    seq, pset, rev = jobs
    _, _, dt_b, A2, _ = packed
    nseq, twoE, ldxz = xz.shape
    E = twoE // 2
    njobs, twoN, ldbc = bc.shape
    h0 = h0.contiguous()
    if cutoff_log2 is None:
        cutoff_log2 = default_cutoff_log2(xz.dtype)
    a = _lib.ScanFixupArgs(_ptr(xz), _ptr(delta), _ptr(bc), _ptr(out), _ptr(dt_b), _ptr(A2), _ptr(seq), _ptr(pset),
                           _ptr(rev), _ptr(h0), L, E, twoN // 2, ldxz, delta.stride(1), ldbc, out.stride(1),
                           nseq, njobs, _dt(xz), channels_per_cta, float(cutoff_log2))
    if seg_ctx is not None and "chunk_dtsum" in seg_ctx:
        # `out` came from the time-parallel kernel (one zero-carry scan of the whole shard): decay the shard's carry-in to the start
        # of every segment (cad_shard_seg_carry) and fix all segments up in parallel — the serial form walks the slow channels through
        # the whole shard with a few warps per SM, which is what bounded the 8-way split
        nseg = seg_ctx["nseg"]
        carry = torch.empty(njobs, nseg, E, twoN // 2, device=xz.device, dtype=torch.float32)
        _lib.check(lib.cad_shard_seg_carry(_ptr(seg_ctx["chunk_dtsum"]), _ptr(A2), _ptr(pset), _ptr(rev), _ptr(h0), _ptr(carry),
                                           njobs, E, seg_ctx["nchunks"], nseg, seg_ctx["per512"], float(cutoff_log2), _stream()),
                   "cad_shard_seg_carry")
        _launched()
        a.nseg, a.seg_carry, a.seg_first = nseg, _ptr(carry), 1
    elif seg_ctx is not None:
        # `out` came from scan variant 20 (every segment scanned from zero): compose the carry of every segment from the shard's
        # carry-in h0 and the segment end states, and fix up ALL segments (the first one included) in this one launch
        nseg = seg_ctx["nseg"]
        carry = torch.empty(njobs, nseg, E, twoN // 2, device=xz.device, dtype=torch.float32)
        _lib.check(lib.cad_seg_carry(_ptr(seg_ctx["seg_state"]), _ptr(seg_ctx["seg_dtsum"]), _ptr(A2), _ptr(pset), _ptr(h0),
                                     _ptr(carry), None, None, njobs, nseg, E, _stream()), "cad_seg_carry")
        _launched()
        if nseg > 1:
            a.nseg, a.seg_carry, a.seg_first, a.cutoff_log2 = nseg, _ptr(carry), 1, float(seg_ctx["cutoff_log2"])
        # nseg == 1: carry[:, 0] == h0 and the whole-sequence mode below is exactly what is needed
    _lib.check(lib.cad_bimamba_scan_fixup(C.byref(a), _stream()), "cad_bimamba_scan_fixup")
    _launched()
    return out


# =====================================================================================================
# fused (RC-equivariant) LM head + masked cross-entropy (csrc/head_ce.cu)
# =====================================================================================================
class _HeadCeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, labels, cmap, loss_weights, ignore_index):
        _require_cuda(hidden, weight, labels)
        lib = _lib.load()
        width = hidden.shape[-1]
        V, D = weight.shape
        rcps = cmap is not None
        h2, rows, ldh = _rows(hidden, width)
        w = weight.to(h2.dtype).contiguous()
        y = labels.reshape(-1).contiguous().long()
        lw = None if loss_weights is None else loss_weights.reshape(-1).contiguous().float()
        nb = lib.cad_head_ce_blocks(rows)
        part = torch.empty(2, nb, device=h2.device, dtype=torch.float32)
        lse = torch.empty(rows, device=h2.device, dtype=torch.float32)
        a = _lib.HeadCeArgs(_ptr(h2), _ptr(w), _ptr(cmap), _ptr(y), _ptr(lw), _ptr(part[0]), _ptr(part[1]), _ptr(lse), None, None,
                            None, rows, D, V, width, ldh, 0, int(ignore_index), int(rcps), _dt(h2), nb)
        _lib.check(lib.cad_head_ce_fwd(C.byref(a), _stream()), "cad_head_ce_fwd")
        _launched()
        sums = part.sum(dim=1)                              # (loss sum, weight sum): deterministic two-stage reduction
        ctx.save_for_backward(h2, w, y, cmap, lw, lse, sums)
        ctx.meta = (rows, D, V, width, ldh, int(ignore_index), rcps, nb, hidden.shape, weight.dtype)
        return sums[0] / sums[1]

    @staticmethod
    def backward(ctx, dloss):
        h2, w, y, cmap, lw, lse, sums = ctx.saved_tensors
        rows, D, V, width, ldh, ignore_index, rcps, nb, hshape, wdtype = ctx.meta
        lib = _lib.load()
        scale = (dloss.float() / sums[1]).reshape(1).contiguous()
        dh = torch.empty(rows, width, device=h2.device, dtype=h2.dtype)
        dwp = torch.empty(nb, V, width, device=h2.device, dtype=torch.float32)
        a = _lib.HeadCeArgs(_ptr(h2), _ptr(w), _ptr(cmap), _ptr(y), _ptr(lw), None, None, _ptr(lse), _ptr(scale), _ptr(dh), _ptr(dwp),
                            rows, D, V, width, ldh, width, ignore_index, int(rcps), _dt(h2), nb)
        _lib.check(lib.cad_head_ce_bwd(C.byref(a), _stream()), "cad_head_ce_bwd")
        _launched()
        dwcat = dwp.sum(dim=0)                              # (V, width)
        dw = dwcat[:, :D]
        if rcps:        # Wcat[v, D + c] = W[cmap[v], D - 1 - c]  ->  dW[cmap[v], D - 1 - c] += dWcat[v, D + c]
            dw = dw.index_add(0, cmap, dwcat[:, D:].flip(-1))
        return dh.view(hshape), dw.to(wdtype), None, None, None, None


def lm_head_cross_entropy(hidden, weight, labels, cmap=None, loss_weights=None, ignore_index=-100):
    """Masked-LM loss straight from the final hidden states: mean (or weight-normalised) cross-entropy of
    `hidden @ Wcat^T` against `labels` over the rows with label != ignore_index, without materialising the logits.
    hidden (..., D) for Caduceus-Ph, (..., 2 D) with cmap (V,) for the RC-equivariant head; weight (V, D)."""
    return _HeadCeFn.apply(hidden, weight, labels, cmap, loss_weights, ignore_index)


# =====================================================================================================
# sequence sharding over NVLink peer memory (csrc/peer_exchange.cu): no library collective on the data path
# =====================================================================================================
def peer_ws_bytes(world, nseq_max, njobs_max, E, N):
    n = _lib.load().cad_peer_ws_bytes(world, nseq_max, njobs_max, E, N)
    if n <= 0:
        raise RuntimeError("cad_peer_ws_bytes: bad geometry")
    return int(n)


def peer_halo_exchange(pctx, xz, L, jobs):
    """Conv halo of a sequence shard through the peers' memory: halo (njobs, E, 3) in xz's dtype.  pctx: _lib.PeerCtx."""
    lib = _lib.load()
    seq, _, rev = jobs
    nseq, twoE, ldxz = xz.shape
    njobs = seq.numel()
    halo = torch.empty(njobs, twoE // 2, 3, device=xz.device, dtype=xz.dtype)
    _lib.check(lib.cad_peer_halo_exchange(C.byref(pctx), _ptr(xz), ldxz, L, nseq, njobs, _ptr(seq), _ptr(rev), _ptr(halo),
                                          _dt(xz), _stream()), "cad_peer_halo_exchange")
    _launched()
    return halo


def peer_carry_exchange(pctx, hlast, dtsum, A2, jobs, want_dtsum_all=False):
    """Boundary states through the peers' memory + carry composition: h0 (njobs, E, N) [, dtsum_all (world, njobs, E)]."""
    lib = _lib.load()
    _, pset, rev = jobs
    njobs, E, N = hlast.shape
    h0 = torch.empty_like(hlast)
    dt_all = torch.empty(pctx.world, njobs, E, device=hlast.device, dtype=torch.float32) if want_dtsum_all else None
    _lib.check(lib.cad_peer_carry_exchange(C.byref(pctx), _ptr(hlast.contiguous()), _ptr(dtsum.contiguous()), _ptr(A2), _ptr(pset),
                                           _ptr(rev), njobs, _ptr(h0), _ptr(dt_all), _stream()), "cad_peer_carry_exchange")
    _launched(2)
    return (h0, dt_all) if want_dtsum_all else h0


def scan_adjoint(xz, delta, bc, dout, packed, jobs, L, cutoff_log2=-40.0, channels_per_cta=0):
    """Dh (njobs, E, N): gradient w.r.t. the shard's carry-in state from the shard's own tokens (zero adjoint
    carry-in) — the quantity the ranks all_gather in a sequence-sharded backward.  See csrc/scan_adjoint.cu."""
    lib = _lib.load()
    seq, pset, rev = jobs
    _, _, dt_b, A2, _ = packed
    nseq, twoE, ldxz = xz.shape
    E = twoE // 2
    njobs, twoN, ldbc = bc.shape
    dout = dout if dout.stride(-1) == 1 and dout.stride(1) % 16 == 0 else dout.contiguous()
    dh = torch.zeros(njobs, E, twoN // 2, device=xz.device, dtype=torch.float32)
    if L == 0:
        return dh
    a = _lib.ScanAdjointArgs(_ptr(xz), _ptr(delta), _ptr(bc), _ptr(dout), _ptr(dt_b), _ptr(A2), _ptr(seq), _ptr(pset),
                             _ptr(rev), _ptr(dh), L, E, twoN // 2, ldxz, delta.stride(1), ldbc, dout.stride(1),
                             nseq, njobs, _dt(xz), channels_per_cta, float(cutoff_log2))
    _lib.check(lib.cad_bimamba_scan_adjoint(C.byref(a), _stream()), "cad_bimamba_scan_adjoint")
    _launched()
    return dh


def conv_halo_grad(xz, du_total, halo, conv_w4, conv_b, jobs, L):
    """Gradient w.r.t. the conv halo (njobs, E, 3), i.e. w.r.t. the 3 `x` samples of the logically PREVIOUS shard
    that this shard's first three conv outputs read.  Tiny (3 tokens per job and channel), so plain torch:
        c[tau] = b + sum_k w[k] xin[tau+k],  xin = [halo, x_logical[0:3]],  dc = du * silu'(c),  tau = 0..2
        dhalo[j] = sum_{tau <= j} w[j - tau] * dc[tau]"""
    seq, pset, rev = jobs
    E = xz.shape[1] // 2
    seq_l, pset_l = seq.long(), pset.long()
    is_rev = rev.to(torch.bool)[:, None, None]
    xs = xz[:, :E]
    x_first = xs[..., 0:3].index_select(0, seq_l)
    x_last = xs[..., L - 3:L].index_select(0, seq_l).flip(-1)
    xin = torch.cat([halo.float(), torch.where(is_rev, x_last, x_first).float()], dim=-1)       # logical -3 .. 2
    du3 = torch.where(is_rev, du_total[..., L - 3:L].flip(-1), du_total[..., 0:3]).float()       # logical 0 .. 2
    w = conv_w4.index_select(0, pset_l).float()                                                    # (njobs, E, 4)
    b = conv_b.index_select(0, pset_l).float()[..., None]
    c = b + sum(w[..., k:k + 1] * xin[..., k:k + 3] for k in range(4))                             # (njobs, E, 3)
    sg = torch.sigmoid(c)
    dc = du3 * sg * (1.0 + c * (1.0 - sg))
    return torch.stack([sum(w[..., j - t] * dc[..., t] for t in range(j + 1)) for j in range(3)], dim=-1)


def scan_bwd(xz, delta, bc, dout, packed, jobs, L, cstate, *, halo=None, h0=None, want_dh0=False,
             channels_per_cta=0, dhlast=None):
    """Backward of scan_fwd: returns dz, du (scan path), ddelta (io dtype), dbc (fp32), ddt_b, dA2, dD, dh0.
    `dhlast` (njobs, E, N): gradient w.r.t. the end state, i.e. the adjoint entering from the next shard."""
    lib = _lib.load()
    seq, pset, rev = jobs
    conv_w4, conv_b, dt_b, A2, Dk = packed
    nseq, twoE, ldxz = xz.shape
    E = twoE // 2
    njobs, twoN, ldbc = bc.shape
    P, _, N = A2.shape
    dev = xz.device
    dout = dout if dout.stride(-1) == 1 and dout.stride(1) % 16 == 0 else dout.contiguous()
    dz = torch.empty(njobs, E, ldxz, device=dev, dtype=xz.dtype)
    du = torch.empty(njobs, E, ldxz, device=dev, dtype=xz.dtype)
    ddelta = torch.empty(njobs, E, ldxz, device=dev, dtype=xz.dtype)
    dbc = torch.zeros(njobs, twoN, ldbc, device=dev, dtype=torch.float32)
    ddt_b = torch.zeros(P, E, device=dev, dtype=torch.float32)
    dA2 = torch.zeros(P, E, N, device=dev, dtype=torch.float32)
    dD = torch.zeros(P, E, device=dev, dtype=torch.float32)
    dh0 = torch.empty(njobs, E, N, device=dev, dtype=torch.float32) if want_dh0 else None
    dhlast = None if dhlast is None else dhlast.float().contiguous()
    a = _lib.ScanBwdArgs(
        _ptr(xz), _ptr(delta), _ptr(bc), _ptr(dout), _ptr(conv_w4), _ptr(conv_b), _ptr(dt_b), _ptr(A2), _ptr(Dk),
        _ptr(seq), _ptr(pset), _ptr(rev), _ptr(halo), _ptr(h0), _ptr(cstate),
        _ptr(dz), _ptr(du), _ptr(ddelta), _ptr(dbc), _ptr(ddt_b), _ptr(dA2), _ptr(dD), _ptr(dh0),
        _ptr(dhlast),
        L, E, N, 4, ldxz, delta.stride(1), ldbc, dout.stride(1), ldxz, ldxz, ldxz,
        nseq, njobs, P, _dt(xz), channels_per_cta)
    _lib.check(lib.cad_bimamba_scan_bwd(C.byref(a), _stream()), "cad_bimamba_scan_bwd")
    _launched()
    return dz, du, ddelta, dbc, ddt_b, dA2, dD, dh0


def conv_silu_bwd(xz, du, conv_w4, conv_b, jobs, L, halo=None):
    """Backward of conv_silu: dx (njobs, E, ld) in the io dtype, dconv_w (P, E, 4), dconv_b (P, E) fp32."""
    lib = _lib.load()
    seq, pset, rev = jobs
    nseq, twoE, ld = xz.shape
    E = twoE // 2
    njobs = seq.numel()
    dx = torch.empty(njobs, E, ld, device=xz.device, dtype=xz.dtype)
    dw = torch.zeros_like(conv_w4)
    db = torch.zeros_like(conv_b)
    a = _lib.ConvBwdArgs(_ptr(xz), _ptr(du), _ptr(dx), _ptr(conv_w4), _ptr(conv_b), _ptr(dw), _ptr(db),
                         _ptr(seq), _ptr(pset), _ptr(rev), _ptr(halo), L, E, ld, du.stride(1), ld,
                         nseq, njobs, _dt(xz))
    _lib.check(lib.cad_conv_silu_bwd(C.byref(a), _stream()), "cad_conv_silu_bwd")
    _launched()
    return dx, dw, db


class _BiMambaCoreFn(torch.autograd.Function):
    """xz -> conv+SiLU -> x_proj -> (dt_proj, B, C) -> fused scan -> gated y, for all jobs of a BiMamba call.
    Differentiable w.r.t. xz, the stacked x_proj / dt_proj weights and the packed scan parameters
    (the backward of upstream's MambaInnerFn, SURVEY.md row A16, minus the in/out projections which stay in autograd)."""

    @staticmethod
    def forward(ctx, xz, w_x, w_dt, conv_w4, conv_b, dt_b, A2, Dk, jobs, L, shard=None):
        packed = (conv_w4, conv_b, dt_b, A2, Dk)
        N = A2.shape[-1]
        E = xz.shape[1] // 2
        pset_l = jobs[1].long()
        sharded = shard is not None and shard.world > 1
        halo = h0 = dt_all = None
        if sharded:
            from . import seqshard
            halo = seqshard.gather_halo(xz[:, :E, :], L, jobs[0], jobs[2], shard).to(xz.dtype)
        u = conv_silu(xz, conv_w4, conv_b, jobs, L, halo=halo)
        xdbl = torch.bmm(w_x.index_select(0, pset_l), u)
        del u
        delta, bc = project_dt_bc(xdbl, w_dt.index_select(0, pset_l), L, N)
        if sharded:
            # TRUE carry for training: zero-carry state pass -> one all_gather -> full scan from h0, so that the saved
            # chunk states (and hence the backward's recomputation) are those of the unsharded sequence
            _, hl, ds, _ = scan_fwd(xz, delta, bc, packed, jobs, L, halo=halo, state_only=True)
            h0, dt_all = seqshard.gather_carry(hl, ds, A2, jobs[1], jobs[2], shard, return_dtsum=True)
        yg, _, _, cstate = scan_fwd(xz, delta, bc, packed, jobs, L, halo=halo, h0=h0, want_chunk_state=True)
        ctx.save_for_backward(xz, w_x, w_dt, conv_w4, conv_b, dt_b, A2, Dk, delta, bc, cstate, xdbl, halo, h0, dt_all)
        ctx.jobs, ctx.L, ctx.shard = jobs, L, (shard if sharded else None)
        return yg

    @staticmethod
    def backward(ctx, dyg):
        xz, w_x, w_dt, conv_w4, conv_b, dt_b, A2, Dk, delta, bc, cstate, xdbl, halo, h0, dt_all = ctx.saved_tensors
        jobs, L, shard = ctx.jobs, ctx.L, ctx.shard
        packed = (conv_w4, conv_b, dt_b, A2, Dk)
        P, _, N = A2.shape
        R = w_dt.shape[-1]
        pset_l, seq_l = jobs[1].long(), jobs[0].long()
        act = xz.dtype
        dhlast = None
        if shard is not None:
            # adjoint carry: local Dh -> one all_gather -> compose the adjoint entering from the next shards
            from . import seqshard
            dh = scan_adjoint(xz, delta, bc, dyg, packed, jobs, L)
            dhlast = seqshard.gather_adjoint(dh, dt_all, A2, jobs[1], jobs[2], shard)
        dz, du, ddelta, dbc, ddt_b, dA2, dD, _ = scan_bwd(xz, delta, bc, dyg, packed, jobs, L, cstate, halo=halo,
                                                          h0=h0, dhlast=dhlast)
        ld = xz.shape[-1]
        # dt_proj and x_proj backward (cuBLAS): d x_dbl = [W_dt^T d dt_raw ; dB ; dC]
        wdt_job = w_dt.index_select(0, pset_l)
        wx_job = w_x.index_select(0, pset_l)
        dxdbl = torch.zeros(xdbl.shape, device=xz.device, dtype=act)
        dxdbl[:, :R, :L] = torch.bmm(wdt_job.transpose(1, 2), ddelta[..., :L])
        dxdbl[:, R:, :L] = dbc[..., :L].to(act)
        dw_dt = torch.zeros(w_dt.shape, device=xz.device, dtype=torch.float32)
        dw_dt.index_add_(0, pset_l, torch.bmm(ddelta[..., :L], xdbl[:, :R, :L].transpose(1, 2)).float())
        u = conv_silu(xz, conv_w4, conv_b, jobs, L, halo=halo)
        dw_x = torch.zeros(w_x.shape, device=xz.device, dtype=torch.float32)
        dw_x.index_add_(0, pset_l, torch.bmm(dxdbl[..., :L], u[..., :L].transpose(1, 2)).float())
        del u
        du_total = torch.baddbmm(du, wx_job.transpose(1, 2), dxdbl)
        dx, dconv_w, dconv_b = conv_silu_bwd(xz, du_total, conv_w4, conv_b, jobs, L, halo=halo)
        if shard is not None:       # the successor's first conv outputs read our edge samples: add their gradient
            dhalo = conv_halo_grad(xz, du_total, halo, conv_w4, conv_b, jobs, L)
            seqshard.exchange_halo_grad(dhalo, dx, L, jobs[0], jobs[2], shard)
        # several jobs (directions) may share one in-proj output: sum their gradients
        E = dx.shape[1]
        nseq, njobs = xz.shape[0], dx.shape[0]
        if njobs == nseq:                                   # one job per in-proj output (untied or unidirectional)
            dxz = torch.cat([dx, dz], dim=1)
        elif njobs % nseq == 0 and _JOB_ADJACENT.get(jobs[0].data_ptr(), False):
            k = njobs // nseq                               # tied projections: the k directions of a sequence are adjacent
            dxz = torch.cat([dx.view(nseq, k, E, ld).sum(1), dz.view(nseq, k, E, ld).sum(1)], dim=1)
        else:
            dxz = torch.zeros(xz.shape, device=xz.device, dtype=act)
            dxz[:, :E].index_add_(0, seq_l, dx)
            dxz[:, E:].index_add_(0, seq_l, dz)
        dxz[..., L:] = 0
        return (dxz, dw_x.to(w_x.dtype), dw_dt.to(w_dt.dtype), dconv_w, dconv_b, ddt_b, dA2, dD, None, None, None)


def bimamba_core(xz, w_x, w_dt, packed, jobs, L, shard=None):
    return _BiMambaCoreFn.apply(xz, w_x, w_dt, *packed, jobs, L, shard)


def microbench(which):
    lib = _lib.load()
    v = C.c_double(0.0)
    _lib.check(lib.cad_microbench(int(which), C.byref(v), _stream()), "cad_microbench")
    return v.value
