"""Idefics2 configuration (reference mlx_vlm/models/idefics2/config.py:8-65: same fields and defaults)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import List, Optional

from ..base import BaseModelConfig


@dataclass
class VisionConfig(BaseModelConfig):
    model_type: str = "idefics2"
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    num_channels: int = 3
    image_size: int = 224
    patch_size: int = 32
    layer_norm_eps: float = 1e-6


@dataclass
class TextConfig(BaseModelConfig):
    model_type: str = "mistral"
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    rms_norm_eps: float = 1e-5
    vocab_size: int = 32003
    rope_theta: float = 1000000.0
    rope_traditional: bool = False
    max_position_embeddings: int = 32768
    tie_word_embeddings: bool = False

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads


@dataclass
class PerceiverConfig(BaseModelConfig):
    model_type: str = "idefics2"
    num_key_value_heads: int = 4
    resampler_depth: int = 3
    resampler_head_dim: int = 96
    resampler_n_heads: int = 16
    resampler_n_latents: int = 64


@dataclass
class ModelConfig(BaseModelConfig):
    text_config: TextConfig
    vision_config: VisionConfig
    perceiver_config: PerceiverConfig
    model_type: str = "idefics2"
    ignore_index: int = -100
    image_token_id: int = 32001
    vocab_size: int = 151936
    image_token_index: Optional[int] = None
    eos_token_id: Optional[List[int]] = None

    def __post_init__(self):
        if self.image_token_index is None:
            self.image_token_index = self.image_token_id

    @classmethod
    def from_dict(cls, params):
        params = dict(params)
        for key, sub in (("text_config", TextConfig), ("vision_config", VisionConfig),
                         ("perceiver_config", PerceiverConfig)):
            if isinstance(params.get(key), dict):
                params[key] = sub.from_dict(params[key])
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


def idefics2_8b_config() -> ModelConfig:
    """Idefics2-8B (SigLIP-SO400M + perceiver + Mistral-7B; BASELINE config 4)."""
    return ModelConfig(text_config=TextConfig(),
                       vision_config=VisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                                                  num_attention_heads=16, image_size=980, patch_size=14),
                       perceiver_config=PerceiverConfig(), vocab_size=32003)
