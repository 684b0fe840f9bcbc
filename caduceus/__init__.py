"""Drop-in alias: `import caduceus` resolves to the B200 implementation, so the reference's trainer
(`registry.model["caduceus_lm"] = "caduceus.modeling_caduceus.CaduceusForMaskedLM"`, ref:src/utils/registry.py:29;
`_target_: caduceus.configuration_caduceus.CaduceusConfig`, ref:configs/model/caduceus.yaml:4) runs unchanged."""
from caduceus_b200 import CaduceusConfig, Caduceus, CaduceusForMaskedLM, CaduceusForSequenceClassification, CaduceusTokenizer  # noqa: F401
