import os
import sys

import pytest

# the virtual-rank tests of tests/test_gpu_multi.py run kernels that WAIT for kernels of other streams: every stream needs its own
# hardware queue (default: 8 connections shared by all streams).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and every kernel must be resident before a waiting kernel runs: with lazy module loading (the CUDA 12 default) the FIRST launch
# of a kernel may synchronise the context, which deadlocks against a kernel of another virtual rank that spins until that launch
# has happened (CUDA programming guide, "Lazy Loading": concurrent execution is not guaranteed across a first launch).  Separate
# processes — the production layout, one rank per process — have separate contexts and are not affected.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

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
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden_loader():
    return golden


# tolerances of the reference's own tests (ref:caduceus/tests/test_rcps.py:34-36): (rtol, atol) per dtype
TOL = {"float32": (6e-4, 2e-3), "float16": (3e-3, 5e-3), "bfloat16": (3e-2, 5e-2)}


def tol(dtype):
    return TOL[str(dtype).replace("torch.", "")]
