"""SmolVLM configuration (reference mlx_vlm/models/smolvlm/config.py:8-80: same fields, defaults and derived values)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass, field
from typing import List, Optional

from ..base import BaseModelConfig


@dataclass
class TextConfig(BaseModelConfig):
    model_type: str = "smolvlm"
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_attention_heads: Optional[int] = None
    rms_norm_eps: float = 1e-5
    vocab_size: int = 49152
    num_key_value_heads: Optional[int] = None
    head_dim: Optional[int] = None
    rope_theta: float = 1000000.0
    num_hidden_layers: int = 32
    rope_traditional: bool = False
    max_position_embeddings: int = 4096
    tie_word_embeddings: bool = False

    def __post_init__(self):
        if self.num_attention_heads is None:
            self.num_attention_heads = self.hidden_size // self.head_dim if self.head_dim else 32
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads


_TOWER_MLP = {768: 3072, 1152: 4304}


@dataclass
class VisionConfig(BaseModelConfig):
    model_type: str = "siglip_vision_model"
    hidden_size: Optional[int] = None
    num_attention_heads: Optional[int] = None
    patch_size: int = 14
    num_hidden_layers: Optional[int] = None
    intermediate_size: Optional[int] = None
    image_size: int = 384
    num_channels: int = 3
    layer_norm_eps: float = 1e-6

    def __post_init__(self):
        if self.hidden_size is None:
            self.hidden_size = 1152
        if self.num_attention_heads is None:
            self.num_attention_heads = self.hidden_size // 64 if self.hidden_size % 64 == 0 else 16
        if self.num_hidden_layers is None:
            self.num_hidden_layers = 12 if self.hidden_size <= 768 else 27
        if self.intermediate_size is None:
            self.intermediate_size = _TOWER_MLP.get(self.hidden_size, self.hidden_size * 4)


@dataclass
class ModelConfig(BaseModelConfig):
    text_config: TextConfig = field(default_factory=TextConfig)
    vision_config: VisionConfig = field(default_factory=VisionConfig)
    model_type: str = "smolvlm"
    ignore_index: int = -100
    vocab_size: int = 49152
    scale_factor: int = 2
    image_token_id: int = 49153
    image_token_index: Optional[int] = None
    eos_token_id: Optional[List[int]] = None

    def __post_init__(self):
        if self.image_token_index is None:
            self.image_token_index = self.image_token_id

    @classmethod
    def from_dict(cls, params):
        params = dict(params)
        for key, sub in (("text_config", TextConfig), ("vision_config", VisionConfig)):
            if isinstance(params.get(key), dict):
                params[key] = sub.from_dict(params[key])
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})
