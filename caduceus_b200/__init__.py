"""caduceus_b200 — B200-native (sm_100a) implementation of Caduceus' bidirectional selective-SSM hot path,
behind the reference's own Python surface (ref:caduceus/__init__.py:5-7)."""
from .configuration_caduceus import CaduceusConfig
from .modeling_caduceus import (BiMambaWrapper, Caduceus, CaduceusForMaskedLM, CaduceusForSequenceClassification,
                                CaduceusMixerModel, CaduceusPreTrainedModel, create_block)
from .tokenization_caduceus import CaduceusTokenizer

__all__ = ["CaduceusConfig", "Caduceus", "CaduceusForMaskedLM", "CaduceusForSequenceClassification",
           "CaduceusTokenizer", "BiMambaWrapper", "CaduceusMixerModel", "CaduceusPreTrainedModel", "create_block",
           "register_auto_classes"]


def register_auto_classes():
    """Make `AutoConfig / AutoModel / AutoModelForMaskedLM / AutoModelForSequenceClassification` resolve
    model_type "caduceus" to this implementation (the reference relies on trust_remote_code, ref:README.md:27-47)."""
    from transformers import (AutoConfig, AutoModel, AutoModelForMaskedLM, AutoModelForSequenceClassification)
    try:
        AutoConfig.register("caduceus", CaduceusConfig)
    except ValueError:
        pass
    for auto, cls in ((AutoModel, Caduceus), (AutoModelForMaskedLM, CaduceusForMaskedLM),
                      (AutoModelForSequenceClassification, CaduceusForSequenceClassification)):
        try:
            auto.register(CaduceusConfig, cls)
        except ValueError:
            pass


register_auto_classes()
