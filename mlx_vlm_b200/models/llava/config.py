"""LLaVA-1.5 configuration (reference mlx_vlm/models/llava/config.py:8-61: same fields and defaults)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Union

from ..base import BaseModelConfig


@dataclass
class TextConfig(BaseModelConfig):
    model_type: str = "llama"
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    intermediate_size: int = 11008
    num_attention_heads: int = 32
    rms_norm_eps: float = 1e-6
    vocab_size: int = 32000
    num_key_value_heads: Optional[int] = None
    rope_theta: float = 10000
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
class VisionConfig(BaseModelConfig):
    model_type: str = "clip_vision_model"
    num_hidden_layers: int = 24
    hidden_size: int = 1024
    intermediate_size: int = 4096
    num_attention_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    projection_dim: int = 768
    vocab_size: int = 32000
    num_channels: int = 3
    layer_norm_eps: float = 1e-5


@dataclass
class ModelConfig(BaseModelConfig):
    text_config: TextConfig
    vision_config: VisionConfig
    model_type: str = "llava"
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
        import inspect
        return cls(**{k: v for k, v in params.items() if k in inspect.signature(cls).parameters})


def llava_15_7b_config() -> ModelConfig:
    """LLaVA-1.5-7B (CLIP-ViT-L/14-336 + Vicuna-7B; BASELINE config 3)."""
    return ModelConfig(text_config=TextConfig(vocab_size=32064, rms_norm_eps=1e-5), vision_config=VisionConfig(),
                       vocab_size=32064)
