"""ctypes binding of the C-ABI in include/caduceus_b200.h.

The product path has NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(UPSTREAM raised through TORCH_CHECK -> RuntimeError; SURVEY.md §8b "Errors").
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcaduceus_b200.so")

CAD_F32, CAD_F16, CAD_BF16 = 0, 1, 2
ABI_VERSION = 5

_p, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float


class EmbeddingArgs(C.Structure):
    _fields_ = [("ids", _p), ("weight", _p), ("cmap", _p), ("out", _p),
                ("B", _i64), ("L", _i64), ("V", _i64), ("D", _i64), ("rcps", _i32), ("dtype", _i32)]


class EmbeddingBwdArgs(C.Structure):
    _fields_ = [("ids", _p), ("cmap", _p), ("dout", _p), ("dweight", _p),
                ("B", _i64), ("L", _i64), ("V", _i64), ("D", _i64), ("rcps", _i32), ("dtype", _i32)]


class AddNormArgs(C.Structure):
    _fields_ = [("x", _p), ("residual", _p), ("weight", _p), ("bias", _p), ("y", _p), ("res_out", _p),
                ("rstd", _p), ("mean", _p),
                ("rows", _i64), ("D", _i64), ("ldx", _i64), ("ldr", _i64), ("ldy", _i64), ("ldo", _i64),
                ("nhalf", _i32), ("swap", _i32), ("wflip_mask", _i32), ("is_rms", _i32),
                ("xdtype", _i32), ("wdtype", _i32), ("res_in_dtype", _i32), ("res_out_dtype", _i32),
                ("eps", _f32)]


class AddNormBwdArgs(C.Structure):
    _fields_ = [("dy", _p), ("dres_out", _p), ("v", _p), ("weight", _p), ("rstd", _p), ("mean", _p),
                ("dx", _p), ("dweight_partial", _p), ("dbias_partial", _p),
                ("rows", _i64), ("D", _i64), ("lddy", _i64), ("lddr", _i64), ("ldv", _i64), ("lddx", _i64),
                ("nhalf", _i32), ("swap", _i32), ("wflip_mask", _i32), ("is_rms", _i32), ("has_bias", _i32),
                ("dydtype", _i32), ("vdtype", _i32), ("wdtype", _i32), ("dxdtype", _i32), ("drdtype", _i32),
                ("nblocks", _i32)]


class ScanFwdArgs(C.Structure):
    _fields_ = [("xz", _p), ("delta", _p), ("bc", _p), ("out", _p),
                ("conv_w", _p), ("conv_b", _p), ("dt_b", _p), ("A2", _p), ("Dskip", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p),
                ("halo", _p), ("h0", _p), ("hlast", _p), ("dtsum", _p), ("chunk_state", _p),
                ("L", _i64), ("E", _i64), ("N", _i64), ("K", _i64),
                ("ldxz", _i64), ("ldd", _i64), ("ldbc", _i64), ("ldo", _i64),
                ("nseq", _i32), ("njobs", _i32), ("npset", _i32), ("io_dtype", _i32), ("channels_per_cta", _i32),
                ("state_only", _i32), ("variant", _i32),
                ("bcT", _p), ("nseg", _i32), ("seg_state", _p), ("seg_dtsum", _p), ("chunk_dtsum", _p)]


class ScanFixupArgs(C.Structure):
    _fields_ = [("xz", _p), ("delta", _p), ("bc", _p), ("out", _p), ("dt_b", _p), ("A2", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p), ("h0", _p),
                ("L", _i64), ("E", _i64), ("N", _i64), ("ldxz", _i64), ("ldd", _i64), ("ldbc", _i64), ("ldo", _i64),
                ("nseq", _i32), ("njobs", _i32), ("io_dtype", _i32), ("channels_per_cta", _i32),
                ("cutoff_log2", _f32), ("nseg", _i32), ("seg_carry", _p), ("seg_first", _i32)]


class ScanAdjointArgs(C.Structure):
    _fields_ = [("xz", _p), ("delta", _p), ("bc", _p), ("dout", _p), ("dt_b", _p), ("A2", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p), ("dh", _p),
                ("L", _i64), ("E", _i64), ("N", _i64), ("ldxz", _i64), ("ldd", _i64), ("ldbc", _i64), ("ldo", _i64),
                ("nseq", _i32), ("njobs", _i32), ("io_dtype", _i32), ("channels_per_cta", _i32),
                ("cutoff_log2", _f32)]


class ScanBwdArgs(C.Structure):
    _fields_ = [("xz", _p), ("delta", _p), ("bc", _p), ("dout", _p),
                ("conv_w", _p), ("conv_b", _p), ("dt_b", _p), ("A2", _p), ("Dskip", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p),
                ("halo", _p), ("h0", _p), ("chunk_state", _p),
                ("dz", _p), ("du", _p), ("ddelta", _p), ("dbc", _p),
                ("ddt_b", _p), ("dA2", _p), ("dDskip", _p), ("dh0", _p), ("dhlast", _p),
                ("L", _i64), ("E", _i64), ("N", _i64), ("K", _i64),
                ("ldxz", _i64), ("ldd", _i64), ("ldbc", _i64), ("ldo", _i64), ("lddz", _i64), ("lddu", _i64),
                ("lddd", _i64),
                ("nseq", _i32), ("njobs", _i32), ("npset", _i32), ("io_dtype", _i32), ("channels_per_cta", _i32)]


class ConvBwdArgs(C.Structure):
    _fields_ = [("xz", _p), ("du", _p), ("dx", _p), ("conv_w", _p), ("conv_b", _p), ("dconv_w", _p), ("dconv_b", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p), ("halo", _p),
                ("L", _i64), ("E", _i64), ("ldxz", _i64), ("lddu", _i64), ("lddx", _i64),
                ("nseq", _i32), ("njobs", _i32), ("io_dtype", _i32)]


class Hg38BatchArgs(C.Structure):
    _fields_ = [("raw", _p), ("rc_flags", _p), ("char_to_id", _p),
                ("masked", _p), ("replaced", _p), ("random_sel", _p), ("random_words", _p),
                ("data", _p), ("target", _p), ("B", _i64), ("L", _i64),
                ("n_id", _i64), ("pad_id", _i64), ("mask_id", _i64)]


class ConvXprojArgs(C.Structure):
    _fields_ = [("xz", _p), ("w_x", _p), ("w_dt", _p), ("conv_w", _p), ("conv_b", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p), ("halo", _p),
                ("delta", _p), ("bc", _p),
                ("L", _i64), ("E", _i64), ("N", _i64), ("R", _i64), ("ldxz", _i64), ("ldd", _i64), ("ldbc", _i64),
                ("nseq", _i32), ("njobs", _i32), ("io_dtype", _i32), ("bcT", _p), ("ldT", _i64), ("w_x_packed", _p)]


class ConvFwdArgs(C.Structure):
    _fields_ = [("xz", _p), ("u", _p), ("conv_w", _p), ("conv_b", _p),
                ("seq_of_job", _p), ("pset_of_job", _p), ("rev_of_job", _p), ("halo", _p),
                ("L", _i64), ("E", _i64), ("ldxz", _i64), ("ldu", _i64),
                ("nseq", _i32), ("njobs", _i32), ("io_dtype", _i32)]


class HeadCeArgs(C.Structure):
    _fields_ = [("hidden", _p), ("weight", _p), ("cmap", _p), ("labels", _p), ("loss_weights", _p),
                ("loss_partial", _p), ("wsum_partial", _p), ("lse", _p), ("dloss_scale", _p), ("dhidden", _p), ("dwcat_partial", _p),
                ("rows", _i64), ("D", _i64), ("V", _i64), ("width", _i64), ("ldh", _i64), ("lddh", _i64), ("ignore_index", _i64),
                ("rcps", _i32), ("io_dtype", _i32), ("nblocks", _i32)]


class WindowMeanArgs(C.Structure):
    _fields_ = [("hidden", _p), ("variant_idx", _p), ("out", _p),
                ("B", _i64), ("L", _i64), ("C", _i64), ("ldh", _i64), ("c0", _i64), ("ldo", _i64),
                ("lo_half", _i32), ("half", _i32), ("flip_len", _i32), ("flip_ch", _i32), ("io_dtype", _i32)]


class PeerCtx(C.Structure):
    _fields_ = [("peer_ws", _p), ("rank", _i32), ("world", _i32),
                ("nseq_max", _i64), ("njobs_max", _i64), ("E", _i64), ("N", _i64)]


# every symbol include/caduceus_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cad_version": (C.c_int, []),
    "cad_last_error": (C.c_char_p, []),
    "cad_sm_count": (C.c_int, []),
    "cad_embedding_fwd": (C.c_int, [C.POINTER(EmbeddingArgs), _p]),
    "cad_embedding_bwd": (C.c_int, [C.POINTER(EmbeddingBwdArgs), _p]),
    "cad_add_norm_fwd": (C.c_int, [C.POINTER(AddNormArgs), _p]),
    "cad_add_norm_bwd": (C.c_int, [C.POINTER(AddNormBwdArgs), _p]),
    "cad_add_norm_bwd_blocks": (C.c_int, [_i64]),
    "cad_bimamba_scan_fwd": (C.c_int, [C.POINTER(ScanFwdArgs), _p]),
    "cad_scan_chunk_len": (C.c_int, []),
    "cad_conv_silu_fwd": (C.c_int, [C.POINTER(ConvFwdArgs), _p]),
    "cad_bimamba_scan_bwd": (C.c_int, [C.POINTER(ScanBwdArgs), _p]),
    "cad_bimamba_scan_fixup": (C.c_int, [C.POINTER(ScanFixupArgs), _p]),
    "cad_conv_silu_bwd": (C.c_int, [C.POINTER(ConvBwdArgs), _p]),
    "cad_conv_xproj_fwd": (C.c_int, [C.POINTER(ConvXprojArgs), _p]),
    "cad_bimamba_scan_adjoint": (C.c_int, [C.POINTER(ScanAdjointArgs), _p]),
    "cad_bc_transpose": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _p]),
    "cad_seg_carry": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p]),
    "cad_shard_seg_carry": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _f32, _p]),
    "cad_peer_ws_bytes": (_i64, [_i32, _i64, _i64, _i64, _i64]),
    "cad_peer_halo_exchange": (C.c_int, [C.POINTER(PeerCtx), _p, _i64, _i64, _i32, _i32, _p, _p, _p, _i32, _p]),
    "cad_peer_carry_exchange": (C.c_int, [C.POINTER(PeerCtx), _p, _p, _p, _p, _p, _i32, _p, _p, _p]),
    "cad_head_ce_blocks": (C.c_int, [_i64]),
    "cad_head_ce_fwd": (C.c_int, [C.POINTER(HeadCeArgs), _p]),
    "cad_head_ce_bwd": (C.c_int, [C.POINTER(HeadCeArgs), _p]),
    "cad_window_mean": (C.c_int, [C.POINTER(WindowMeanArgs), _p]),
    "cad_hg38_batch_fwd": (C.c_int, [C.POINTER(Hg38BatchArgs), _p]),
    "cad_microbench": (C.c_int, [C.c_int, C.POINTER(C.c_double), _p]),
}

_lock = threading.Lock()
_lib = None


def load():
    """dlopen the in-tree library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"caduceus_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m caduceus_b200.build` "
                "(or __graft_entry__.build()). There is no CPU / eager fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)     # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.cad_version() != ABI_VERSION:
            raise RuntimeError(f"caduceus_b200: ABI mismatch (library {lib.cad_version()}, binding {ABI_VERSION}); rebuild")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().cad_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
