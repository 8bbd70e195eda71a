"""Qwen2.5-VL configuration: the schema of reference mlx_vlm/models/qwen2_5_vl/config.py:8-90 as tables; validation of the
M-RoPE `rope_scaling` and the root-level language-model parameters copied into `text_config` as in the reference."""
from __future__ import annotations

import inspect

from ..config_schema import config_class

_VISION = """
    model_type              str         'qwen2_5_vl'
    depth                   int         32
    hidden_size             int         1280
    intermediate_size       int         3420
    out_hidden_size         int         1536
    num_heads               int         16
    image_size              int         384
    patch_size              int         14
    vocab_size              int         32000
    mlp_ratio               float       4.0
    in_channels             int         3
    layer_norm_eps          float       1e-6
    spatial_patch_size      int         14
    spatial_merge_size      int         2
    tokens_per_second       int         2
    temporal_patch_size     int         2
    window_size             int         112
    fullatt_block_indexes   List[int]   [7, 15, 23, 31]
"""
_TEXT = """
    model_type               str                                           -
    hidden_size              int                                           -
    num_hidden_layers        int                                           -
    intermediate_size        int                                           -
    num_attention_heads      int                                           -
    rms_norm_eps             float                                         -
    vocab_size               int                                           -
    num_key_value_heads      Optional[int]                                 None
    max_position_embeddings  Optional[int]                                 128000
    rope_theta               float                                         1000000.0
    rope_traditional         bool                                          False
    rope_scaling             Optional[Dict[str,Union[float,str,list]]]     None
    tie_word_embeddings      bool                                          True
"""
_MODEL = """
    text_config                      object                -
    vision_config                    object                -
    model_type                       str                   -
    ignore_index                     int                   -100
    image_token_id                   int                   151655
    video_token_id                   int                   151656
    vision_start_token_id            int                   151652
    vision_end_token_id              int                   151653
    vision_token_id                  int                   151654
    vision_feature_select_strategy   str                   'default'
    vision_feature_layer             int                   -2
    vocab_size                       int                   32000
    eos_token_id                     Optional[List[int]]   None
"""


def _text_rules(self):
    if self.num_key_value_heads is None:
        self.num_key_value_heads = self.num_attention_heads
    scaling = self.rope_scaling
    if scaling:
        if not {"mrope_section", "type"} <= set(scaling):
            raise ValueError("rope_scaling must contain keys {'mrope_section', 'type'}")
        if scaling["type"] not in ("mrope", "default"):
            raise ValueError("rope_scaling type must be 'mrope' or 'default'")


def _mrope_section(self):
    """sections of the rotary half-dimension per position axis (default: rope_utils.py:1030-1040)"""
    return list((self.rope_scaling or {}).get("mrope_section") or (24, 20, 20))


def _from_root(cls, params):
    """config.json keeps the language-model parameters at the root: everything except `vision_config` is the text config"""
    params = dict(params)
    params["text_config"] = {k: v for k, v in params.items() if k != "vision_config"}
    kw = {k: v for k, v in params.items() if k in inspect.signature(cls).parameters}
    for key, sub in (("text_config", TextConfig), ("vision_config", VisionConfig)):
        if isinstance(kw.get(key), dict):
            kw[key] = sub.from_dict(kw[key])
    return cls(**kw)


VisionConfig = config_class("VisionConfig", __name__, _VISION)
TextConfig = config_class("TextConfig", __name__, _TEXT, _text_rules, {"mrope_section": property(_mrope_section)})
ModelConfig = config_class("ModelConfig", __name__, _MODEL, members={"from_dict": classmethod(_from_root)})


def qwen2_5_vl_3b_config() -> "ModelConfig":
    """Qwen2.5-VL-3B-Instruct dims (HF config.json of the checkpoint)"""
    text = TextConfig(model_type="qwen2_5_vl", hidden_size=2048, num_hidden_layers=36, intermediate_size=11008,
                      num_attention_heads=16, rms_norm_eps=1e-6, vocab_size=151936, num_key_value_heads=2,
                      rope_scaling={"type": "mrope", "mrope_section": [16, 24, 24]}, tie_word_embeddings=True)
    return ModelConfig(text_config=text, vision_config=VisionConfig(out_hidden_size=2048), model_type="qwen2_5_vl",
                       vocab_size=151936, eos_token_id=[151645, 151643])
