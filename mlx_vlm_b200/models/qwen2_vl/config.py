"""Qwen2-VL configs — same dataclasses/fields as reference
mlx_vlm/models/qwen2_vl/config.py:9-86 (ModelConfig.from_dict copies the root-level
text parameters into text_config the same way)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass
from typing import Dict, List, Optional, Union

from ..base import BaseModelConfig


@dataclass
class VisionConfig(BaseModelConfig):
    model_type: str = "qwen2_vl"
    depth: int = 32
    embed_dim: int = 1280
    hidden_size: int = 1536
    num_heads: int = 16
    image_size: int = 384
    patch_size: int = 14
    vocab_size: int = 32000
    mlp_ratio: float = 4.0
    in_channels: int = 3
    layer_norm_eps: float = 1e-6
    spatial_patch_size: int = 14
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2


@dataclass
class TextConfig(BaseModelConfig):
    model_type: str
    hidden_size: int
    num_hidden_layers: int
    intermediate_size: int
    num_attention_heads: int
    rms_norm_eps: float
    vocab_size: int
    num_key_value_heads: Optional[int] = 8
    max_position_embeddings: Optional[int] = 40960
    rope_theta: float = 1000000.0
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Union[float, str, list]]] = None
    tie_word_embeddings: bool = False
    sliding_window: int = 32768
    use_sliding_window: bool = False
    use_cache: bool = True

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.rope_scaling:
            required_keys = {"mrope_section", "type"}
            if not all(key in self.rope_scaling for key in required_keys):
                raise ValueError(f"rope_scaling must contain keys {required_keys}")
            if not self.rope_scaling["type"] in ["mrope", "default"]:
                raise ValueError("rope_scaling type must be 'mrope' or 'default'")

    @property
    def mrope_section(self):
        # get_mrope_section default (rope_utils.py:1030-1040)
        return list((self.rope_scaling or {}).get("mrope_section") or (24, 20, 20))


@dataclass
class ModelConfig(BaseModelConfig):
    text_config: TextConfig
    vision_config: VisionConfig
    model_type: str
    ignore_index: int = -100
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_feature_select_strategy: str = "default"
    vision_feature_layer: int = -2
    vocab_size: int = 32000
    eos_token_id: Optional[List[int]] = None

    @classmethod
    def from_dict(cls, params):
        params = dict(params)
        excluded_keys = {"vision_config"}
        text = dict(filter(lambda x: x[0] not in excluded_keys, params.items()))
        params["text_config"] = text
        kw = {k: v for k, v in params.items() if k in inspect.signature(cls).parameters}
        if isinstance(kw.get("text_config"), dict):
            kw["text_config"] = TextConfig.from_dict(kw["text_config"])
        if isinstance(kw.get("vision_config"), dict):
            kw["vision_config"] = VisionConfig.from_dict(kw["vision_config"])
        return cls(**kw)


def qwen2_vl_2b_config() -> ModelConfig:
    """Qwen2-VL-2B-Instruct dims (HF config.json of the checkpoint)."""
    text = TextConfig(
        model_type="qwen2_vl", hidden_size=1536, num_hidden_layers=28, intermediate_size=8960,
        num_attention_heads=12, rms_norm_eps=1e-6, vocab_size=151936, num_key_value_heads=2,
        max_position_embeddings=32768, rope_theta=1000000.0,
        rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, tie_word_embeddings=True)
    vision = VisionConfig(depth=32, embed_dim=1280, hidden_size=1536, num_heads=16)
    return ModelConfig(text_config=text, vision_config=vision, model_type="qwen2_vl",
                       vocab_size=151936, eos_token_id=[151645, 151643])


def qwen2_vl_7b_config() -> ModelConfig:
    text = TextConfig(
        model_type="qwen2_vl", hidden_size=3584, num_hidden_layers=28, intermediate_size=18944,
        num_attention_heads=28, rms_norm_eps=1e-6, vocab_size=152064, num_key_value_heads=4,
        max_position_embeddings=32768, rope_theta=1000000.0,
        rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, tie_word_embeddings=False)
    vision = VisionConfig(depth=32, embed_dim=1280, hidden_size=3584, num_heads=16)
    return ModelConfig(text_config=text, vision_config=vision, model_type="qwen2_vl",
                       vocab_size=152064, eos_token_id=[151645, 151643])
