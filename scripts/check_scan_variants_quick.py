"""Quick GPU check (no pytest start-up): the forward-scan variants against the float64 boundary restatement and
against each other.  python scripts/check_scan_variants_quick.py"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import test_gpu_scan_variants as T  # noqa: E402

for L in (17, 513, 2300):
    for rev in (0, 1):
        T.test_v7_no_replay_variant_vs_boundary_restatement(L, rev)
        print("v7 boundary ok", L, rev, flush=True)
T.test_v7_state_outputs_match_v3_fp32()
print("v7 state outputs ok", flush=True)
T.test_v4_agrees_with_v3_on_identical_inputs()
print("v4 vs v3 ok", flush=True)
