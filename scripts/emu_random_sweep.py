"""Random sweep of the emulated scan kernels (CPU only; test infrastructure, not a product path):
    python scripts/emu_random_sweep.py [--seconds 300] [--seed 1]
Draws (length, job table, CTA shape, hooks, variant) at random and runs the SAME checks as tests/test_emu_scan_*.py on them:
forward variants 4 and 9..12, backward variant 2, and variant 20's whole pipeline (segment scans -> carries -> segment-mode
fix-up), plain and as one shard of a longer sequence.  The fixed test cases cover the geometry classes one can think of; this
finds the ones one did not (it found the double-entered conv halo of variant 20 at ragged ends with < 3 masked tokens)."""
import argparse
import ctypes as C
import os
import random
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from caduceus_b200 import _lib  # noqa: E402
import test_emu_scan_bwd_v2 as TB  # noqa: E402
import test_emu_scan_fixup as TF  # noqa: E402
import test_emu_scan_v4 as T4  # noqa: E402
import test_emu_scan_v9 as T9  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=300.0)
ap.add_argument("--seed", type=int, default=1)
args = ap.parse_args()

so = os.path.join(ROOT, "tests", "emu", "libemu_scan.so")
if not os.path.exists(so):
    sys.exit("build tests/emu/libemu_scan.so first: python -m pytest tests/test_emu_scan_v4.py -q")
lib = C.CDLL(so)
for name, argt in (("emu_scan_v4", [C.POINTER(_lib.ScanFwdArgs), C.c_int]),
                   ("emu_scan_v9", [C.POINTER(_lib.ScanFwdArgs), C.c_int, C.c_int, C.c_int]),
                   ("emu_scan_v20", [C.POINTER(_lib.ScanFwdArgs), C.c_int]),
                   ("emu_scan_fixup", [C.POINTER(_lib.ScanFixupArgs), C.c_int]),
                   ("emu_scan_bwd_v2", [C.POINTER(_lib.ScanBwdArgs), C.c_int])):
    getattr(lib, name).restype = C.c_int
    getattr(lib, name).argtypes = argt

rng = random.Random(args.seed)
LS = [1, 2, 3, 6, 7, 8, 9, 15, 16, 17, 31, 33, 63, 255, 256, 257, 500, 511, 512, 513, 527, 528, 529, 767, 1000, 1022, 1023, 1024, 1025,
      1040, 1300, 1500, 1535, 1537, 1800, 2047, 2049, 2600]
SPECS = [[(0, 0, 0)], [(0, 0, 1)], [(0, 0, 0), (0, 1, 1)], [(0, 0, 1), (0, 1, 0), (1, 0, 0), (1, 1, 1)]]
count = {"v4": 0, "v9..12": 0, "bwd2": 0, "v20": 0, "v20 shard": 0}
t0 = time.time()
while time.time() - t0 < args.seconds:
    L, spec, G, E = rng.choice(LS), rng.choice(SPECS), rng.choice([1, 2, 3]), rng.choice([2, 3, 4, 6])
    half = rng.choice([torch.bfloat16, torch.float16])
    r, cfg = rng.random(), None
    try:
        if r < 0.35:
            tile32, pipe, hooks = rng.random() < 0.6, rng.choice([0, 1]), rng.random() < 0.6
            state_only = hooks and rng.random() < 0.25
            dt_ready = (not hooks) and rng.random() < 0.2
            cfg = ("v9..12", L, spec, G, E, tile32, pipe, hooks, state_only, dt_ready)
            T9._run(lib, L, E=E, spec=spec, dtype=half, G=G, seed=rng.randrange(10 ** 6), pipe=pipe, hooks=hooks and not state_only,
                    state_only=state_only, tile32=tile32, dt_ready=dt_ready)
            count["v9..12"] += 1
        elif r < 0.6:
            hooks = rng.random() < 0.6
            cfg = ("bwd2", L, spec, G, E, hooks)
            TB._run(lib, min(L, 1100), E=E, spec=spec[:2], dtype=rng.choice([torch.float32, torch.bfloat16]), G=G,
                    seed=rng.randrange(10 ** 6), hooks=hooks)
            count["bwd2"] += 1
        elif r < 0.7:
            cfg = ("v4", L, spec, G, E + E % 2)
            T4._run_emu(lib, L, E + E % 2, spec, half, G, rng.randrange(10 ** 6))
            count["v4"] += 1
        else:
            shard, nseg = rng.random() < 0.5, rng.choice([2, 3, 4, 5, 7, 9])
            E20, W20, var = rng.choice([8, 32, 40, 64, 70]), rng.choice([1, 2, 3]), rng.choice([20, 20, 21, 22, 23])
            cfg = ("v20 shard" if shard else "v20", L, nseg, E20, W20, var)
            TF._pipeline(lib, L, nseg, rng.choice([-24.0, -40.0]), shard=shard, E=E20, W=W20, variant=var)
            count["v20 shard" if shard else "v20"] += 1
    except AssertionError as exc:
        print("FAIL", cfg, str(exc)[:300], flush=True)
        sys.exit(1)
print("ok", count, f"in {time.time() - t0:.0f} s")
