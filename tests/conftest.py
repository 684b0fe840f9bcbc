import os
import sys

import pytest

# tests/test_gpu_multi.py runs kernels that WAIT for kernels of other streams (virtual ranks, in subprocesses of their own): every
# stream needs its own hardware queue (default: 8 connections shared by all streams).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, ORACLE):          # ORACLE on the path makes `mamba_ssm` (the CPU shim) importable — tests only
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=True)


@pytest.fixture(scope="session")
def golden_loader():
    return golden


# tolerances of the reference's own tests (ref:caduceus/tests/test_rcps.py:34-36): (rtol, atol) per dtype
TOL = {"float32": (6e-4, 2e-3), "float16": (3e-3, 5e-3), "bfloat16": (3e-2, 5e-2)}


def tol(dtype):
    return TOL[str(dtype).replace("torch.", "")]
