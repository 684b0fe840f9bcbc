"""`CaduceusConfig` — field-for-field the configuration surface of ref:caduceus/configuration_caduceus.py:10-55
(model_type "caduceus"), so HF `AutoConfig` and the reference's hydra `_target_`
(ref:configs/model/caduceus.yaml:4) resolve to an object the B200 implementation understands.
"""
from transformers import PretrainedConfig

# (name, default) in the reference's order: the first eight mirror upstream MambaConfig, `norm_epsilon` is
# create_block's layer-norm epsilon, `initializer_cfg` feeds _init_weights, the rest are Caduceus-specific.
_FIELDS = (
    ("d_model", 2560),
    ("n_layer", 64),
    ("vocab_size", 50277),
    ("ssm_cfg", None),
    ("rms_norm", True),
    ("residual_in_fp32", True),
    ("fused_add_norm", True),
    ("pad_vocab_size_multiple", 8),
    ("norm_epsilon", 1e-5),
    ("initializer_cfg", None),
    ("bidirectional", True),
    ("bidirectional_strategy", "add"),
    ("bidirectional_weight_tie", True),
    ("rcps", False),
    ("complement_map", None),     # token id -> complement id; consumed by RCPSEmbedding / RCPSLMHead
)


class CaduceusConfig(PretrainedConfig):
    model_type = "caduceus"

    def __init__(self, d_model=2560, n_layer=64, vocab_size=50277, ssm_cfg=None, rms_norm=True,
                 residual_in_fp32=True, fused_add_norm=True, pad_vocab_size_multiple=8, norm_epsilon=1e-5,
                 initializer_cfg=None, bidirectional=True, bidirectional_strategy="add",
                 bidirectional_weight_tie=True, rcps=False, complement_map=None, **kwargs):
        # transformers 4.38 (the reference's pin) tied lm_head to the embedding unless told otherwise;
        # transformers 5 dropped that default, so state it.
        kwargs.setdefault("tie_word_embeddings", True)
        super().__init__(**kwargs)
        given = locals()
        for name, _default in _FIELDS:
            setattr(self, name, given[name])
        if self.complement_map is not None:
            # JSON round trips turn the int keys into strings; consumers rely on insertion order of the VALUES
            # (ref:caduceus/modeling_rcps.py:31-34), which survives.
            self.complement_map = {int(k): int(v) for k, v in self.complement_map.items()}
