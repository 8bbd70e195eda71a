"""LLaVA-Next configuration (reference mlx_vlm/models/llava_next/config.py:8-60: same fields and defaults — the
Mistral-7B language model: 8 kv heads, rope_theta 1e6, intermediate 14336, vocabulary 32064)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

from ..base import BaseModelConfig
from ..llava.config import VisionConfig  # noqa: F401  (identical fields and defaults)


@dataclass
class TextConfig(BaseModelConfig):
    model_type: str = "mistral"
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    intermediate_size: int = 14336
    num_attention_heads: int = 32
    rms_norm_eps: float = 1e-05
    vocab_size: int = 32064
    num_key_value_heads: Optional[int] = 8
    rope_theta: float = 1000000
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Union[float, str]]] = None
    max_position_embeddings: int = 4096
    tie_word_embeddings: bool = False

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.rope_scaling:
            need = {"factor", "type"}
            if not all(k in self.rope_scaling for k in need):
                raise ValueError(f"rope_scaling must contain keys {need}")
            if self.rope_scaling["type"] != "linear":
                raise ValueError("rope_scaling 'type' currently only supports 'linear'")


@dataclass
class ModelConfig(BaseModelConfig):
    text_config: TextConfig
    vision_config: VisionConfig
    model_type: str = "llava_next"
    ignore_index: int = -100
    image_token_index: int = 32000
    vision_feature_select_strategy: str = "default"
    vision_feature_layer: int = -2
    vocab_size: int = 32000
    eos_token_id: Optional[List[int]] = None

    @classmethod
    def from_dict(cls, params):
        params = dict(params)
        for key, sub in (("text_config", TextConfig), ("vision_config", VisionConfig)):
            if isinstance(params.get(key), dict):
                params[key] = sub.from_dict(params[key])
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})
