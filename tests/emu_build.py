"""Builds tests/emu/libemu_scan.so — the lane = channel scan kernel source (caduceus_b200/csrc/scan_fwd_v20.cuh) compiled for the
host against the SIMT emulation of tests/emu/simt_emu.h (lanes are threads, shared memory / cp.async / bulk copies / mbarriers are
modelled).  Test infrastructure only: the product runs only the CUDA build of the same source."""
import ctypes as C
import os
import subprocess

import pytest

from caduceus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libemu_scan.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_scan.cpp", "simt_emu.h")] + [
        os.path.join(ROOT, "caduceus_b200", "csrc", "scan_fwd_v20.cuh"), os.path.join(ROOT, "caduceus_b200", "csrc", "simt.cuh"),
        os.path.join(ROOT, "include", "caduceus_b200.h")]
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")):
        pytest.skip("CUDA headers not found")
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-DCAD_EMULATE", "-I", EMU_DIR,
                        "-I", CUDA_INC, srcs[0], "-o", EMU_SO], check=True)
    lib = C.CDLL(EMU_SO)
    lib.emu_scan_v20.restype = C.c_int
    lib.emu_scan_v20.argtypes = [C.POINTER(_lib.ScanFwdArgs), C.c_int]
    return lib
