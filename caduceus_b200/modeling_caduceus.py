"""Caduceus model classes with the reference's public surface (ref:caduceus/modeling_caduceus.py): same class
names, constructor signatures, `state_dict` keys and outputs — `create_block`, `BiMambaWrapper`,
`CaduceusEmbeddings`, `CaduceusMixerModel`, `CaduceusPreTrainedModel`, `Caduceus`, `CaduceusForMaskedLM`,
`CaduceusForSequenceClassification` — running on the sm_100a kernels of this package.

What differs underneath: one in_proj per strand instead of one per direction, no
materialised flips, one fused scan launch for all directions/strands, one out_proj over the concatenated
directions, RC halves handled by reversed weights inside the norm kernel.
"""
import inspect
import math
from functools import partial
from typing import Optional, Tuple, Union

import torch
from torch import nn
from torch.nn import functional as F
from transformers import PreTrainedModel
from transformers.modeling_outputs import BaseModelOutputWithNoAttention, MaskedLMOutput, SequenceClassifierOutput

from . import functional as CF
from .configuration_caduceus import CaduceusConfig
from .modeling_rcps import RCPSAddNormWrapper, RCPSEmbedding, RCPSLMHead, RCPSMambaBlock
from .modules import Block, Mamba, RMSNorm, bimamba_inner, layer_norm_fn, rms_norm_fn  # noqa: F401

_STRATEGIES = ("add", "ew_multiply")


def create_block(d_model, ssm_cfg=None, norm_epsilon=1e-5, rms_norm=False, residual_in_fp32=False,
                 fused_add_norm=False, layer_idx=None, bidirectional=True, bidirectional_strategy="add",
                 bidirectional_weight_tie=True, rcps=False, device=None, dtype=None):
    """One pre-norm residual block: (RCPS)Block(norm, BiMambaWrapper).  ref:caduceus/modeling_caduceus.py:33-84."""
    factory = {"device": device, "dtype": dtype}
    mixer_cls = partial(BiMambaWrapper, layer_idx=layer_idx, **(ssm_cfg or {}), bidirectional=bidirectional,
                        bidirectional_strategy=bidirectional_strategy,
                        bidirectional_weight_tie=bidirectional_weight_tie, **factory)
    norm_cls = partial(RMSNorm if rms_norm else nn.LayerNorm, eps=norm_epsilon, **factory)
    block_cls = RCPSMambaBlock if rcps else Block
    extra = {"mlp_cls": nn.Identity} if "mlp_cls" in inspect.signature(block_cls.__init__).parameters else {}
    block = block_cls(d_model, mixer_cls, norm_cls=norm_cls, fused_add_norm=fused_add_norm,
                      residual_in_fp32=residual_in_fp32, **extra)
    block.layer_idx = layer_idx
    return block


class BiMambaWrapper(nn.Module):
    """Two Mamba parameter sets (`mamba_fwd`, `mamba_rev`), optionally sharing in/out projections, evaluated as
    M_fwd(h) (+ or *) flip(M_rev(flip(h))) — in one fused pipeline.  ref:caduceus/modeling_caduceus.py:87-140."""

    def __init__(self, d_model: int, bidirectional: bool = True, bidirectional_strategy: Optional[str] = "add",
                 bidirectional_weight_tie: bool = True, **mamba_kwargs):
        super().__init__()
        if bidirectional and bidirectional_strategy is None:
            bidirectional_strategy = "add"
        if bidirectional and bidirectional_strategy not in _STRATEGIES:
            raise NotImplementedError(f"`{bidirectional_strategy}` strategy for bi-directionality is not implemented!")
        self.bidirectional = bidirectional
        self.bidirectional_strategy = bidirectional_strategy
        self.bidirectional_weight_tie = bool(bidirectional and bidirectional_weight_tie)
        self.mamba_fwd = Mamba(d_model=d_model, **mamba_kwargs)
        self.mamba_rev = None
        if bidirectional:
            self.mamba_rev = Mamba(d_model=d_model, **mamba_kwargs)
            self.tie_projections()

    def tie_projections(self):
        """mamba_rev shares in_proj / out_proj with mamba_fwd (ref:caduceus/modeling_caduceus.py:114-118): the two
        projections hold most of the parameters; conv / x_proj / dt_proj / A / D stay per direction."""
        if self.bidirectional_weight_tie:
            tied = {}
            for proj in ("in_proj", "out_proj"):
                src, dst = getattr(self.mamba_fwd, proj), getattr(self.mamba_rev, proj)
                dst.weight = src.weight
                dst.bias = src.bias
                tied[f"mamba_rev.{proj}.weight"] = f"mamba_fwd.{proj}.weight"
                if src.bias is not None:
                    tied[f"mamba_rev.{proj}.bias"] = f"mamba_fwd.{proj}.bias"
            # transformers >= 5 walks submodules for this attribute when it serialises shared tensors
            self._tied_weights_keys = tied

    def forward(self, hidden_states, inference_params=None):
        """hidden_states (B, L, D) -> (B, L, D)."""
        if inference_params is not None:
            raise NotImplementedError("caduceus_b200: step-wise decoding (inference_params) is outside the hot path")
        return bimamba_inner(hidden_states, self.mamba_fwd, self.mamba_rev, self.bidirectional_strategy, nstrand=1)

    def forward_rcps(self, hidden_states, inference_params=None):
        """(B, L, 2D) -> (B, L, 2D): cat[self(x1), rc(self(rc(x2)))] without flips (used by RCPSWrapper)."""
        if inference_params is not None:
            raise NotImplementedError("caduceus_b200: step-wise decoding (inference_params) is outside the hot path")
        return bimamba_inner(hidden_states, self.mamba_fwd, self.mamba_rev, self.bidirectional_strategy, nstrand=2)


class CaduceusEmbeddings(nn.Module):
    def __init__(self, config: CaduceusConfig, device=None, dtype=None):
        super().__init__()
        factory = {"device": device, "dtype": dtype}
        if config.rcps:
            self.word_embeddings = RCPSEmbedding(config.vocab_size, config.d_model, config.complement_map, **factory)
        else:
            self.word_embeddings = nn.Embedding(config.vocab_size, config.d_model, **factory)

    def forward(self, input_ids):
        """input_ids (B, L) -> (B, L, D) or (B, L, 2D)."""
        if isinstance(self.word_embeddings, RCPSEmbedding):
            return self.word_embeddings(input_ids)
        if isinstance(self.word_embeddings, nn.Embedding) and self.word_embeddings.padding_idx is None \
                and self.word_embeddings.max_norm is None:
            return CF.embedding(input_ids, self.word_embeddings.weight)
        return self.word_embeddings(input_ids)      # user-replaced embedding module


class CaduceusMixerModel(nn.Module):
    def __init__(self, config: CaduceusConfig, device=None, dtype=None) -> None:
        super().__init__()
        factory = {"device": device, "dtype": dtype}
        self.fused_add_norm = config.fused_add_norm
        self.rcps = config.rcps
        self.residual_in_fp32 = config.residual_in_fp32
        self.embeddings = CaduceusEmbeddings(config, **factory)
        self.layers = nn.ModuleList([
            create_block(config.d_model, ssm_cfg=config.ssm_cfg, norm_epsilon=config.norm_epsilon,
                         rms_norm=config.rms_norm, residual_in_fp32=config.residual_in_fp32,
                         fused_add_norm=config.fused_add_norm, layer_idx=i, bidirectional=config.bidirectional,
                         bidirectional_strategy=config.bidirectional_strategy,
                         bidirectional_weight_tie=config.bidirectional_weight_tie, rcps=config.rcps, **factory)
            for i in range(config.n_layer)])
        norm_f = (RMSNorm if config.rms_norm else nn.LayerNorm)(config.d_model, eps=config.norm_epsilon, **factory)
        self.norm_f = norm_f if (config.fused_add_norm or not config.rcps) else RCPSAddNormWrapper(norm_f)

    def forward(self, input_ids, inputs_embeds=None, output_hidden_states=False):
        all_hidden_states = []
        hidden_states = inputs_embeds if inputs_embeds is not None else self.embeddings(input_ids)
        residual = None
        for layer in self.layers:
            if output_hidden_states:
                all_hidden_states.append(hidden_states)
            hidden_states, residual = layer(hidden_states, residual, inference_params=None)

        # final add + norm (ref:caduceus/modeling_caduceus.py:233-275); the residual is not needed afterwards
        if not self.fused_add_norm:
            if self.rcps:
                hidden_states = self.norm_f(hidden_states, residual=residual, prenorm=False)
            else:
                residual = (hidden_states + residual) if residual is not None else hidden_states
                hidden_states = CF.add_norm(residual.to(dtype=self.norm_f.weight.dtype), self.norm_f.weight,
                                            self.norm_f.bias, eps=self.norm_f.eps,
                                            is_rms=isinstance(self.norm_f, RMSNorm))
        else:
            # PS: halves normalised separately, RC half with the reversed weight, and — unlike the fused block —
            # NO half swap here (ref:caduceus/modeling_caduceus.py:244-262)
            hidden_states = CF.add_norm(
                hidden_states, self.norm_f.weight, self.norm_f.bias, residual=residual, eps=self.norm_f.eps,
                is_rms=isinstance(self.norm_f, RMSNorm), prenorm=False, residual_in_fp32=self.residual_in_fp32,
                nhalf=2 if self.rcps else 1, swap=0, wflip_mask=2 if self.rcps else 0)
            if output_hidden_states:
                all_hidden_states.append(hidden_states)
        return hidden_states, all_hidden_states


def cross_entropy(logits, y, ignore_index=-100):
    return F.cross_entropy(logits.view(-1, logits.shape[-1]), y.view(-1), ignore_index=ignore_index)


def weighted_cross_entropy(logits, y, loss_weights, ignore_index=-100):
    """Per-token weighted CE, weights renormalised over the non-ignored tokens
    (ref:caduceus/modeling_caduceus.py:286-294; note: like the reference it zeroes `loss_weights` in place)."""
    y = y.view(-1)
    ce = F.cross_entropy(logits.view(-1, logits.shape[-1]), y, ignore_index=ignore_index, reduction="none")
    loss_weights = loss_weights.view(-1)
    loss_weights[y == ignore_index] = 0.0
    return (ce * (loss_weights / loss_weights.sum())).sum()


class CaduceusPreTrainedModel(PreTrainedModel):
    config_class = CaduceusConfig
    base_model_prefix = "caduceus"
    supports_gradient_checkpointing = False
    _no_split_modules = ["BiMambaWrapper"]

    def get_expanded_tied_weights_keys(self, all_submodels: bool = False) -> dict:
        """{duplicate parameter name: canonical name} for every parameter this model shares: the BiMamba projection
        ties (structural, independent of `config.tie_word_embeddings`) plus the head/embedding tie.  transformers >= 5
        needs the map to save / reload shared tensors; 4.x simply dropped duplicates with a warning."""
        try:
            mapping = dict(super().get_expanded_tied_weights_keys(all_submodels=all_submodels))
        except (AttributeError, TypeError):        # transformers 4.x has no such hook
            mapping = {}
        seen = {}
        for name, p in self.named_parameters(remove_duplicate=False):
            if id(p) in seen:
                mapping.setdefault(name, seen[id(p)])
            else:
                seen[id(p)] = name
        return mapping

    def _retie_structural(self):
        for m in self.modules():
            if isinstance(m, BiMambaWrapper):
                m.tie_projections()

    def _init_weights(self, module, initializer_range=0.02, **kwargs):
        """Mamba's GPT-2-style init (ref:caduceus/modeling_caduceus.py:304-341): zero Linear biases (except
        `_no_reinit` ones such as dt_proj.bias), N(0, range) embeddings, out_proj rescaled by
        1/sqrt(n_residuals_per_layer * n_layer)."""
        cfg = self.config.initializer_cfg or {}

        def loaded(p):      # transformers >= 5 flags parameters it has just read from a checkpoint: leave those alone
            return getattr(p, "_is_hf_initialized", False)

        if isinstance(module, nn.Linear):
            if module.bias is not None and not getattr(module.bias, "_no_reinit", False) and not loaded(module.bias):
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.Embedding):
            if not loaded(module.weight):
                nn.init.normal_(module.weight, std=cfg.get("initializer_range", initializer_range))
        if cfg.get("rescale_prenorm_residual", True):
            scale = math.sqrt(cfg.get("n_residuals_per_layer", 1) * self.config.n_layer)
            for name, p in module.named_parameters():
                if name in ("out_proj.weight", "fc2.weight") and not loaded(p):
                    # re-draw before scaling so that repeated calls do not shrink the weight repeatedly
                    nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                    with torch.no_grad():
                        p /= scale


def _return_dict(config, return_dict):
    if return_dict is not None:
        return return_dict
    return getattr(config, "return_dict", True)


class Caduceus(CaduceusPreTrainedModel):
    """Backbone: embeddings -> n_layer blocks -> final norm.  ref:caduceus/modeling_caduceus.py:344-389."""

    def tie_weights(self, **kwargs):
        self._retie_structural()

    def __init__(self, config: CaduceusConfig, device=None, dtype=None, **kwargs):
        super().__init__(config)
        if config.rcps and config.complement_map is None:
            raise AssertionError("Complement map must be provided for RCPS.")
        # vocabulary padding mutates the config, and the complement map is extended with identities
        # (ref:caduceus/modeling_caduceus.py:353-357)
        rem = config.vocab_size % config.pad_vocab_size_multiple
        if rem != 0:
            config.vocab_size += config.pad_vocab_size_multiple - rem
        if config.complement_map is not None:
            for i in range(len(config.complement_map), config.vocab_size):
                config.complement_map[i] = i
        self.config = config
        self.backbone = CaduceusMixerModel(config, device=device, dtype=dtype, **kwargs)
        # the reference leaves the bare backbone un-"post_init"-ed (it is initialised through the head models);
        # transformers >= 5 needs the bookkeeping post_init() sets up to load it standalone via AutoModel
        self.post_init()

    def forward(self, input_ids: torch.LongTensor = None, inputs_embeds: Optional[torch.FloatTensor] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None,
                ) -> Union[torch.Tensor, Tuple, BaseModelOutputWithNoAttention]:
        if output_hidden_states is None:
            output_hidden_states = self.config.output_hidden_states
        return_dict = _return_dict(self.config, return_dict)
        hidden_states, all_hidden_states = self.backbone(input_ids, inputs_embeds=inputs_embeds,
                                                         output_hidden_states=output_hidden_states)
        if return_dict:
            return BaseModelOutputWithNoAttention(last_hidden_state=hidden_states,
                                                  hidden_states=all_hidden_states if output_hidden_states else None)
        if output_hidden_states:
            return hidden_states, all_hidden_states
        return hidden_states


class CaduceusForMaskedLM(CaduceusPreTrainedModel):
    """Backbone + LM head.  ref:caduceus/modeling_caduceus.py:392-492."""

    def __init__(self, config: CaduceusConfig, device=None, dtype=None, **kwargs):
        super().__init__(config, **kwargs)
        self.caduceus = Caduceus(config, device=device, dtype=dtype, **kwargs)
        if config.rcps:
            # sizes are read after Caduceus() because it may have padded the vocabulary
            self.lm_head = RCPSLMHead(complement_map=self.config.complement_map, vocab_size=self.config.vocab_size,
                                      true_dim=config.d_model, dtype=dtype)
        else:
            self.lm_head = nn.Linear(config.d_model, self.config.vocab_size, bias=False, device=device, dtype=dtype)
        # head <-> table tie, declared for transformers >= 5 (PS always; Ph when config.tie_word_embeddings)
        emb = "caduceus.backbone.embeddings.word_embeddings."
        if config.rcps:
            self._tied_weights_keys = {"lm_head.lm_head.weight": emb + "embedding.weight"}
        elif getattr(config, "tie_word_embeddings", True):
            self._tied_weights_keys = {"lm_head.weight": emb + "weight"}
        self.post_init()

    def get_input_embeddings(self):
        return self.caduceus.backbone.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        if self.config.rcps:
            raise NotImplementedError("Setting input embeddings for RCPS LM is not supported.")
        self.caduceus.backbone.embeddings.word_embeddings = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, new_embeddings):
        if self.config.rcps:
            raise NotImplementedError("Setting output embeddings for RCPS LM is not supported.")
        self.lm_head = new_embeddings

    def tie_weights(self, **kwargs):
        """PS always shares the table with the head (ref:caduceus/modeling_caduceus.py:434-439); Ph follows
        `config.tie_word_embeddings` (True under the reference's transformers pin)."""
        self._retie_structural()
        if self.config.rcps:
            self.lm_head.set_weight(self.get_input_embeddings().weight)
        elif getattr(self.config, "tie_word_embeddings", True):
            self.lm_head.weight = self.get_input_embeddings().weight

    def get_decoder(self):
        return self.caduceus

    def set_decoder(self, decoder):
        self.caduceus = decoder

    def forward(self, input_ids: torch.LongTensor = None, inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None, loss_weights: Optional[torch.FloatTensor] = None,
                output_hidden_states: Optional[bool] = None, return_dict: Optional[bool] = None,
                ) -> Union[Tuple, MaskedLMOutput]:
        if output_hidden_states is None:
            output_hidden_states = self.config.output_hidden_states
        return_dict = _return_dict(self.config, return_dict)
        outputs = self.caduceus(input_ids=input_ids, inputs_embeds=inputs_embeds,
                                output_hidden_states=output_hidden_states, return_dict=return_dict)
        hidden_states = outputs[0] if isinstance(outputs, (tuple, BaseModelOutputWithNoAttention)) else outputs
        if labels is not None and getattr(self.config, "fused_head_loss", False) and hidden_states.is_cuda:
            # opt-in (config.fused_head_loss = True; not a reference field): the masked-LM loss straight from the hidden states —
            # head GEMM and cross-entropy only at the non-ignored positions, fp32 (B, L, V) logits never materialised
            # (csrc/head_ce.cu).  `logits` of the output is None in this mode.
            return self._fused_loss_output(hidden_states, labels, loss_weights, outputs, return_dict)
        logits = self.lm_head(hidden_states).float()

        loss = None
        if labels is not None:
            ignore = getattr(self.config, "pad_token_id", None)
            ignore = -100 if ignore is None else ignore
            if loss_weights is not None:
                loss = weighted_cross_entropy(logits, labels, loss_weights, ignore_index=ignore)
            else:
                loss = cross_entropy(logits, labels, ignore_index=ignore)

        if not return_dict:
            rest = tuple(outputs[1:]) if isinstance(outputs, tuple) else ()
            output = (logits,) + rest
            return (loss,) + output if loss is not None else output
        return MaskedLMOutput(loss=loss, logits=logits, hidden_states=outputs.hidden_states)


def _fused_loss_output(self, hidden_states, labels, loss_weights, outputs, return_dict):
    ignore = getattr(self.config, "pad_token_id", None)
    ignore = -100 if ignore is None else ignore
    head = self.lm_head
    if self.config.rcps:
        if head.lm_head.bias is not None:
            raise NotImplementedError("fused_head_loss: a head bias is not supported")
        weight, cmap = head.weight, head.complement_map
    else:
        if head.bias is not None:
            raise NotImplementedError("fused_head_loss: a head bias is not supported")
        weight, cmap = head.weight, None
    if loss_weights is not None:          # like the reference (ref:caduceus/modeling_caduceus.py:291), in place
        loss_weights.view(-1)[labels.view(-1) == ignore] = 0.0
    loss = CF.lm_head_cross_entropy(hidden_states, weight, labels, cmap=cmap, loss_weights=loss_weights, ignore_index=ignore)
    if not return_dict:
        rest = tuple(outputs[1:]) if isinstance(outputs, tuple) else ()
        return (loss, None) + rest
    return MaskedLMOutput(loss=loss, logits=None, hidden_states=outputs.hidden_states)


CaduceusForMaskedLM._fused_loss_output = _fused_loss_output


class CaduceusForSequenceClassification(CaduceusPreTrainedModel):
    """Pooled classifier with optional RC conjoining.  ref:caduceus/modeling_caduceus.py:495-640."""

    def __init__(self, config: CaduceusConfig, pooling_strategy: str = "mean", conjoin_train: bool = False,
                 conjoin_eval: bool = False, device=None, dtype=None, **kwargs):
        super().__init__(config, **kwargs)
        if pooling_strategy not in ("mean", "max", "first", "last"):
            raise NotImplementedError(f"Pooling strategy `{pooling_strategy}` not implemented.")
        self.pooling_strategy = pooling_strategy
        self.num_labels = kwargs.get("num_labels", config.num_labels)
        self.caduceus = Caduceus(config, device=device, dtype=dtype, **kwargs)
        self.score = nn.Linear(config.d_model, self.num_labels, bias=False)
        self.conjoin_train = conjoin_train
        self.conjoin_eval = conjoin_eval
        self.post_init()
        self.init_scorer()

    def tie_weights(self, **kwargs):
        self._retie_structural()

    def init_scorer(self, initializer_range=0.02):
        cfg = self.config.initializer_cfg or {}
        self.score.weight.data.normal_(std=cfg.get("initializer_range", initializer_range))

    def get_input_embeddings(self):
        return self.caduceus.backbone.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        if self.config.rcps:
            raise NotImplementedError("Setting input embeddings for RCPS LM is not supported.")
        self.caduceus.backbone.embeddings.word_embeddings = value

    def pool_hidden_states(self, hidden_states, sequence_length_dim=1):
        if self.pooling_strategy == "mean":
            return hidden_states.mean(dim=sequence_length_dim)
        if self.pooling_strategy == "max":
            return hidden_states.max(dim=sequence_length_dim).values
        index = -1 if self.pooling_strategy == "last" else 0
        return hidden_states.select(sequence_length_dim, index)

    def forward(self, input_ids: torch.LongTensor = None, inputs_embeds: Optional[torch.FloatTensor] = None,
                labels: Optional[torch.LongTensor] = None, output_hidden_states: Optional[bool] = None,
                return_dict: Optional[bool] = None) -> Union[Tuple, SequenceClassifierOutput]:
        return_dict = _return_dict(self.config, return_dict)
        run = partial(self.caduceus, output_hidden_states=output_hidden_states, return_dict=return_dict)
        d = self.config.d_model
        if self.config.rcps:
            # 2*d_model channels -> stack the forward half and the RC-aligned second half on a new last axis
            outs = run(input_ids, inputs_embeds=inputs_embeds)
            hidden = torch.stack([outs[0][..., :d], torch.flip(outs[0][..., d:], dims=[1, 2])], dim=-1)
        elif self.conjoin_train or (self.conjoin_eval and not self.training):
            if input_ids is None or input_ids.ndim != 3:
                raise AssertionError("`input_ids` must be a 3D tensor (batch, length, strand) for conjoining.")
            outs = run(input_ids[..., 0], inputs_embeds=None)
            outs_rc = run(input_ids[..., 1], inputs_embeds=None)
            hidden = torch.stack([outs[0], outs_rc[0]], dim=-1)
        else:
            outs = run(input_ids, inputs_embeds=None)
            hidden = outs[0]

        pooled = self.pool_hidden_states(hidden)
        if hidden.ndim == 4:       # (batch, length, d_model, 2): score both strands with shared weights, average
            logits = (self.score(pooled[..., 0]) + self.score(pooled[..., 1])) / 2
        else:
            logits = self.score(pooled)

        loss = None
        if labels is not None:
            labels = labels.to(logits.device)
            if self.config.problem_type is None:
                if self.num_labels == 1:
                    self.config.problem_type = "regression"
                elif labels.dtype in (torch.long, torch.int):
                    self.config.problem_type = "single_label_classification"
                else:
                    self.config.problem_type = "multi_label_classification"
            if self.config.problem_type == "regression":
                loss = F.mse_loss(logits.squeeze(), labels.squeeze()) if self.num_labels == 1 \
                    else F.mse_loss(logits, labels)
            elif self.config.problem_type == "single_label_classification":
                loss = F.cross_entropy(logits.view(-1, self.num_labels), labels.view(-1))
            else:
                loss = F.binary_cross_entropy_with_logits(logits, labels)
        if not return_dict:
            output = (logits,) + tuple(outs[1:])
            return ((loss,) + output) if loss is not None else output
        return SequenceClassifierOutput(loss=loss, logits=logits, hidden_states=outs.hidden_states)
