"""Contract types between the generation engine and per-model modules.

Mirrors reference mlx_vlm/models/base.py:54-118 (LanguageModelOutput,
InputEmbeddingsFeatures, BaseModelConfig) — same field names, so
`InputEmbeddingsFeatures.to_dict()` keys flow verbatim into
`language_model(**kwargs)` exactly as in generate/ar.py:399-405.  Arrays are
torch tensors (device memory containers) instead of mx.array.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Any, Dict, List, Optional


@dataclass
class LanguageModelOutput:
    logits: Any
    hidden_states: Optional[List[Any]] = None
    cross_attention_states: Optional[List[Any]] = None
    encoder_outputs: Optional[List[Any]] = None
    gdn_states: Optional[List] = None
    shared_kv_states: Optional[Dict[str, tuple]] = None


@dataclass
class InputEmbeddingsFeatures:
    inputs_embeds: Any
    attention_mask_4d: Optional[Any] = None
    visual_pos_masks: Optional[Any] = None
    deepstack_visual_embeds: Optional[Any] = None
    per_layer_inputs: Optional[Any] = None
    cross_attention_states: Optional[Any] = None
    cross_attention_mask: Optional[Any] = None
    full_text_row_masked_out_mask: Optional[Any] = None
    decoder_inputs_embeds: Optional[Any] = None
    attention_mask: Optional[Any] = None
    position_ids: Optional[Any] = None
    pos_hw: Optional[Any] = None
    rope_deltas: Optional[Any] = None

    def to_dict(self):
        return {
            "inputs_embeds": self.inputs_embeds,
            "attention_mask_4d": self.attention_mask_4d,
            "visual_pos_masks": self.visual_pos_masks,
            "deepstack_visual_embeds": self.deepstack_visual_embeds,
            "per_layer_inputs": self.per_layer_inputs,
            "cross_attention_states": self.cross_attention_states,
            "cross_attention_mask": self.cross_attention_mask,
            "full_text_row_masked_out_mask": self.full_text_row_masked_out_mask,
            "decoder_inputs_embeds": self.decoder_inputs_embeds,
            "attention_mask": self.attention_mask,
            "position_ids": self.position_ids,
            "pos_hw": self.pos_hw,
            "rope_deltas": self.rope_deltas,
        }


@dataclass
class BaseModelConfig:
    @classmethod
    def from_dict(cls, params):
        if not params:
            return cls()
        return cls(**{k: v for k, v in params.items()
                      if k in inspect.signature(cls).parameters})

    def to_dict(self):
        return {k: v for k, v in self.__dict__.items() if v is not None}
