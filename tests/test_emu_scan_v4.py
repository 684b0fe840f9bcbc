"""CPU check of the v4 (paired-channel) scan kernel's index logic through a SIMT emulation.

caduceus_b200/csrc/scan_fwd_v4.cuh is compiled for the host with -DCAD_EMULATE (tests/emu/: lanes are threads,
shuffles are exchanges, TMA / mbarrier are modelled, the 128-byte swizzle is applied) and run on the same argument
block the CUDA kernel takes.  The checker is a float64 restatement of the operator at the kernel boundary
(include/caduceus_b200.h, "Semantics in LOGICAL time tau"; upstream selective_scan_ref + causal_conv1d, SURVEY.md
App. A.1/A.2/A.5).  This is NOT a product path: the product runs only the CUDA build of the same source.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libemu_scan.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_scan.cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_fwd_v4.cuh"),
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_fwd_v9.cuh"),
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_bwd_v2.cuh"),
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_fwd_v20.cuh"),
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_fixup.cuh"), os.path.join(ROOT, "include", "caduceus_b200.h")]
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("CUDA headers not found")
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DCAD_EMULATE", "-I", EMU_DIR,
                        "-I", CUDA_INC, srcs[0], "-o", EMU_SO], check=True)
    lib = C.CDLL(EMU_SO)
    lib.emu_scan_v4.restype = C.c_int
    lib.emu_scan_v4.argtypes = [C.POINTER(_lib.ScanFwdArgs), C.c_int]
    lib.emu_scan_v9.restype = C.c_int
    lib.emu_scan_v9.argtypes = [C.POINTER(_lib.ScanFwdArgs), C.c_int, C.c_int, C.c_int]
    lib.emu_scan_v20.restype = C.c_int
    lib.emu_scan_v20.argtypes = [C.POINTER(_lib.ScanFwdArgs), C.c_int]
    lib.emu_scan_fixup.restype = C.c_int
    lib.emu_scan_fixup.argtypes = [C.POINTER(_lib.ScanFixupArgs), C.c_int]
    lib.emu_scan_bwd_v2.restype = C.c_int
    lib.emu_scan_bwd_v2.argtypes = [C.POINTER(_lib.ScanBwdArgs), C.c_int]
    return lib


from scan_boundary_ref import _problem, boundary_ref  # noqa: E402,F401  (re-exported for test_gpu_scan_v4.py)


def _run_emu(lib, L, E, spec, dtype, G, seed):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, dtype, seed)
    njobs = len(spec)
    out = torch.full((njobs, E, ld), float("nan")).to(dtype)
    p = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    a = _lib.ScanFwdArgs(p(xz), p(delta), p(bc), p(out), p(conv_w4), p(conv_b), p(dt_b), p(A2), p(Dk),
                         p(tabs[0]), p(tabs[1]), p(tabs[2]), None, None, None, None, None,
                         L, E, 16, 4, ld, ld, ldbc, ld, xz.shape[0], njobs, conv_w4.shape[0],
                         _lib.CAD_BF16 if dtype == torch.bfloat16 else _lib.CAD_F16, G, 0, 0, 4, None, 0)
    assert lib.emu_scan_v4(C.byref(a), G) == 0
    f = lambda t: t.float().numpy()   # noqa: E731
    ref = boundary_ref(f(xz), f(delta), f(bc), f(conv_w4), f(conv_b), f(dt_b), f(A2), f(Dk),
                       [s for s, _, _ in spec], [q for _, q, _ in spec], [r for _, _, r in spec], L)
    got = out.float().numpy()
    # pad columns [L, ld) must not be written
    assert np.isnan(got[..., L:]).all(), "kernel wrote into the pad columns"
    got = got[..., :L]
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11      # one rounding of the 16-bit output + fp32 noise
    err = np.abs(got - ref)
    bound = 1e-4 + 1.5 * eps * np.abs(ref)
    assert np.isfinite(got).all()
    assert (err <= bound).all(), f"L={L}: max err {err.max():.3e}, worst excess {(err - bound).max():.3e} at {np.unravel_index((err - bound).argmax(), err.shape)}"


@pytest.mark.parametrize("L", [1, 5, 16, 17, 511, 512, 513, 1030])
@pytest.mark.parametrize("rev", [0, 1])
def test_emulated_v4_ragged_lengths(emu, L, rev):
    _run_emu(emu, L, E=4, spec=[(0, 0, rev)], dtype=torch.bfloat16, G=2, seed=100 + L)


def test_emulated_v4_jobs_psets_and_idle_warps(emu):
    """4 jobs as Caduceus-PS orders them (strand x direction, rev = direction XOR strand), 2 sequences, 2 parameter
    sets, E/2 = 3 pairs over CTAs of 2 warps (the last CTA has an idle warp), three chunks with a ragged tail."""
    spec = [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    _run_emu(emu, 1100, E=6, spec=spec, dtype=torch.bfloat16, G=2, seed=7)


def test_emulated_v4_fp16_many_chunks(emu):
    """five chunks: both TMA tiles are re-armed twice (mbarrier phases 0, 1, 0), fp16 I/O, 3 warps per CTA."""
    _run_emu(emu, 2300, E=6, spec=[(0, 0, 0), (0, 0, 1)], dtype=torch.float16, G=3, seed=9)
