"""Qwen2-VL configuration objects.

The field names, order and defaults are the schema of the reference's config dataclasses
(mlx_vlm/models/qwen2_vl/config.py:9-86) — a checkpoint's `config.json` must load unchanged — so they
are kept as DATA (the tables below) and the classes are generated from them; the behaviour around
them (kv-head default, rope_scaling validation, root-level text parameters copied into `text_config`)
is restated and pinned by tests/test_batch_host.py::test_qwen2_vl_config_schema_and_from_dict.
"""
from __future__ import annotations

import inspect
from dataclasses import field, make_dataclass
from typing import Dict, List, Optional, Union

from ..base import BaseModelConfig

_REQUIRED = object()

# (name, type, default) — the reference's schema, in its order
_VISION_SCHEMA = (
    ("model_type", str, "qwen2_vl"), ("depth", int, 32), ("embed_dim", int, 1280),
    ("hidden_size", int, 1536), ("num_heads", int, 16), ("image_size", int, 384),
    ("patch_size", int, 14), ("vocab_size", int, 32000), ("mlp_ratio", float, 4.0),
    ("in_channels", int, 3), ("layer_norm_eps", float, 1e-6), ("spatial_patch_size", int, 14),
    ("spatial_merge_size", int, 2), ("temporal_patch_size", int, 2),
)
_TEXT_SCHEMA = (
    ("model_type", str, _REQUIRED), ("hidden_size", int, _REQUIRED), ("num_hidden_layers", int, _REQUIRED),
    ("intermediate_size", int, _REQUIRED), ("num_attention_heads", int, _REQUIRED),
    ("rms_norm_eps", float, _REQUIRED), ("vocab_size", int, _REQUIRED),
    ("num_key_value_heads", Optional[int], 8), ("max_position_embeddings", Optional[int], 40960),
    ("rope_theta", float, 1000000.0), ("rope_traditional", bool, False),
    ("rope_scaling", Optional[Dict[str, Union[float, str, list]]], None),
    ("tie_word_embeddings", bool, False), ("sliding_window", int, 32768),
    ("use_sliding_window", bool, False), ("use_cache", bool, True),
)
_MODEL_SCHEMA = (
    ("text_config", object, _REQUIRED), ("vision_config", object, _REQUIRED), ("model_type", str, _REQUIRED),
    ("ignore_index", int, -100), ("image_token_id", int, 151655), ("video_token_id", int, 151656),
    ("vision_start_token_id", int, 151652), ("vision_feature_select_strategy", str, "default"),
    ("vision_feature_layer", int, -2), ("vocab_size", int, 32000), ("eos_token_id", Optional[List[int]], None),
)


def _build(name: str, schema, namespace=None):
    fields = [(n, t) if d is _REQUIRED else (n, t, field(default=d)) for n, t, d in schema]
    cls = make_dataclass(name, fields, bases=(BaseModelConfig,), namespace=namespace or {})
    cls.__module__ = __name__
    return cls


VisionConfig = _build("VisionConfig", _VISION_SCHEMA)


def _text_post_init(self):
    """kv heads default to MHA; an M-RoPE `rope_scaling` must name its sections and a known type."""
    if self.num_key_value_heads is None:
        self.num_key_value_heads = self.num_attention_heads
    scaling = self.rope_scaling
    if scaling:
        missing = {"mrope_section", "type"} - set(scaling)
        if missing:
            raise ValueError(f"rope_scaling must contain keys {{'mrope_section', 'type'}} (missing {sorted(missing)})")
        if scaling["type"] not in ("mrope", "default"):
            raise ValueError("rope_scaling type must be 'mrope' or 'default'")


def _mrope_section(self):
    """Sections of the rotary half-dimension per position axis (t, h, w); the default is
    get_mrope_section's (rope_utils.py:1030-1040)."""
    return list((self.rope_scaling or {}).get("mrope_section") or (24, 20, 20))


TextConfig = _build("TextConfig", _TEXT_SCHEMA,
                    {"__post_init__": _text_post_init, "mrope_section": property(_mrope_section)})


def _model_from_dict(cls, params):
    """`config.json` keeps the language-model parameters at the root: every root key except
    `vision_config` becomes the text config (the reference does the same), then the known keys build
    the dataclass; nested dicts become config objects."""
    params = dict(params)
    params["text_config"] = {k: v for k, v in params.items() if k != "vision_config"}
    known = inspect.signature(cls).parameters
    kw = {k: v for k, v in params.items() if k in known}
    for key, sub in (("text_config", TextConfig), ("vision_config", VisionConfig)):
        if isinstance(kw.get(key), dict):
            kw[key] = sub.from_dict(kw[key])
    return cls(**kw)


ModelConfig = _build("ModelConfig", _MODEL_SCHEMA, {"from_dict": classmethod(_model_from_dict)})


def _preset(hidden, inter, heads, kv, vocab, tied) -> "ModelConfig":
    text = TextConfig(model_type="qwen2_vl", hidden_size=hidden, num_hidden_layers=28, intermediate_size=inter,
                      num_attention_heads=heads, rms_norm_eps=1e-6, vocab_size=vocab, num_key_value_heads=kv,
                      max_position_embeddings=32768, rope_theta=1000000.0,
                      rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, tie_word_embeddings=tied)
    vision = VisionConfig(depth=32, embed_dim=1280, hidden_size=hidden, num_heads=16)
    return ModelConfig(text_config=text, vision_config=vision, model_type="qwen2_vl", vocab_size=vocab,
                       eos_token_id=[151645, 151643])


def qwen2_vl_2b_config() -> "ModelConfig":
    """Qwen2-VL-2B-Instruct dims (HF config.json of the checkpoint)."""
    return _preset(1536, 8960, 12, 2, 151936, True)


def qwen2_vl_7b_config() -> "ModelConfig":
    """Qwen2-VL-7B-Instruct dims (config C5)."""
    return _preset(3584, 18944, 28, 4, 152064, False)
