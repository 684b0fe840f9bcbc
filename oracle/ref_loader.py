"""ORACLE (test infrastructure only) — load the reference's OWN `caduceus/` package on top of the
CPU `mamba_ssm` shim in this directory.

The reference package (ref:caduceus/__init__.py:5-7) is loaded under the module name `ref_caduceus`
so it can coexist with this repo's drop-in `caduceus` package. `/root/reference` exists only in the
build container: callers on the GPU box must use the committed fixtures in tests/golden/ instead.

One compatibility patch is applied (SURVEY.md §8b "compat"): transformers 5.x calls
`tie_weights(recompute_mapping=False)` from `post_init()`, while the reference (written against 4.38.1,
ref:caduceus_env.yml:45) defines `tie_weights(self)` (ref:caduceus/modeling_caduceus.py:434).
"""
import importlib.util
import os
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("CADUCEUS_REFERENCE_DIR", "/root/reference")


def shim_on_path():
    if ORACLE_DIR not in sys.path:
        sys.path.insert(0, ORACLE_DIR)


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "caduceus", "modeling_caduceus.py"))


def load_reference(name="ref_caduceus"):
    """Return the reference package (modules: .configuration_caduceus, .modeling_caduceus,
    .modeling_rcps, .tokenization_caduceus), running on the oracle shim."""
    if name in sys.modules:
        return sys.modules[name]
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    shim_on_path()
    pkg_dir = os.path.join(REFERENCE_DIR, "caduceus")
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules[name] = pkg
    spec.loader.exec_module(pkg)

    mc = sys.modules[f"{name}.modeling_caduceus"]
    orig = mc.CaduceusForMaskedLM.tie_weights

    def tie_weights(self, **kwargs):
        if self.config.rcps:
            return orig(self)
        # transformers 4.38.1 (the reference's pin) tied lm_head to the embedding by default
        if getattr(self.config, "tie_word_embeddings", True):
            self.lm_head.weight = self.get_input_embeddings().weight

    mc.CaduceusForMaskedLM.tie_weights = tie_weights
    return pkg
