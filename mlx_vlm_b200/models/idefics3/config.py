"""Idefics3 configuration: the schema of reference mlx_vlm/models/idefics3/config.py:8-58 as tables (the reference leaves
the sizes of the text / vision configs without defaults; the defaults here only add convenience for hand-built configs)."""
from __future__ import annotations

from ..config_schema import config_class, image_token_alias, kv_heads_default, nested_from_dict

_TEXT = """
    model_type               str             'llama'
    hidden_size              int             4096
    intermediate_size        int             14336
    num_attention_heads      int             32
    rms_norm_eps             float           1e-5
    vocab_size               int             128259
    num_key_value_heads      Optional[int]   8
    rope_theta               float           1000000.0
    num_hidden_layers        int             32
    rope_traditional         bool            False
    max_position_embeddings  int             4096
    tie_word_embeddings      bool            False
"""
_VISION = """
    model_type            str     'idefics3'
    hidden_size           int     1152
    num_attention_heads   int     16
    patch_size            int     14
    num_hidden_layers     int     12
    intermediate_size     int     3072
    image_size            int     224
    num_channels          int     3
    layer_norm_eps        float   1e-6
"""
_MODEL = """
    text_config         object                -
    vision_config       object                -
    model_type          str                   'idefics3'
    ignore_index        int                   -100
    vocab_size          int                   128259
    scale_factor        int                   2
    image_token_id      int                   49153
    image_token_index   Optional[int]         None
    eos_token_id        Optional[List[int]]   None
"""

TextConfig = config_class("TextConfig", __name__, _TEXT, kv_heads_default)
VisionConfig = config_class("VisionConfig", __name__, _VISION)
ModelConfig = config_class("ModelConfig", __name__, _MODEL, image_token_alias,
                           {"from_dict": nested_from_dict(text_config=TextConfig, vision_config=VisionConfig)})


def idefics3_8b_config() -> "ModelConfig":
    """Idefics3-8B-Llama3: SigLIP-SO400M at 364 px (26 x 26 patches, 169 tokens per crop after the 2x2 pixel shuffle) +
    Llama-3.1-8B"""
    return ModelConfig(text_config=TextConfig(rope_theta=500000.0, max_position_embeddings=131072),
                       vision_config=VisionConfig(hidden_size=1152, intermediate_size=4304, num_hidden_layers=27,
                                                  num_attention_heads=16, image_size=364, patch_size=14),
                       image_token_id=128257)
