"""Idefics2 configuration: the schema of reference mlx_vlm/models/idefics2/config.py:8-65 as tables (the reference leaves
`model_type` without a default; the defaults here only add convenience for hand-built configs)."""
from __future__ import annotations

from ..config_schema import config_class, image_token_alias, kv_heads_default, nested_from_dict

_VISION = """
    model_type            str     'idefics2'
    hidden_size           int     4096
    intermediate_size     int     14336
    num_hidden_layers     int     32
    num_attention_heads   int     32
    num_key_value_heads   int     8
    num_channels          int     3
    image_size            int     224
    patch_size            int     32
    layer_norm_eps        float   1e-6
"""
_TEXT = """
    model_type               str     'mistral'
    hidden_size              int     4096
    intermediate_size        int     14336
    num_hidden_layers        int     32
    num_attention_heads      int     32
    num_key_value_heads      int     8
    rms_norm_eps             float   1e-5
    vocab_size               int     32003
    rope_theta               float   1000000.0
    rope_traditional         bool    False
    max_position_embeddings  int     32768
    tie_word_embeddings      bool    False
"""
_PERCEIVER = """
    model_type            str   'idefics2'
    num_key_value_heads   int   4
    resampler_depth       int   3
    resampler_head_dim    int   96
    resampler_n_heads     int   16
    resampler_n_latents   int   64
"""
_MODEL = """
    text_config         object                -
    vision_config       object                -
    perceiver_config    object                -
    model_type          str                   'idefics2'
    ignore_index        int                   -100
    image_token_id      int                   32001
    vocab_size          int                   151936
    image_token_index   Optional[int]         None
    eos_token_id        Optional[List[int]]   None
"""

VisionConfig = config_class("VisionConfig", __name__, _VISION)
TextConfig = config_class("TextConfig", __name__, _TEXT, kv_heads_default)
PerceiverConfig = config_class("PerceiverConfig", __name__, _PERCEIVER)
ModelConfig = config_class("ModelConfig", __name__, _MODEL, image_token_alias,
                           {"from_dict": nested_from_dict(text_config=TextConfig, vision_config=VisionConfig,
                                                          perceiver_config=PerceiverConfig)})


def idefics2_8b_config() -> "ModelConfig":
    """Idefics2-8B (SigLIP-SO400M + perceiver + Mistral-7B; BASELINE config 4)."""
    return ModelConfig(text_config=TextConfig(),
                       vision_config=VisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                                                  num_attention_heads=16, image_size=980, patch_size=14),
                       perceiver_config=PerceiverConfig(), vocab_size=32003)
