"""LLaVA-1.5 configuration: the schema of reference mlx_vlm/models/llava/config.py:8-61 as tables (models/config_schema.py)."""
from __future__ import annotations

from ..config_schema import config_class, nested_from_dict

_TEXT = """
    model_type               str                                      'llama'
    hidden_size              int                                      4096
    num_hidden_layers        int                                      32
    intermediate_size        int                                      11008
    num_attention_heads      int                                      32
    rms_norm_eps             float                                    1e-6
    vocab_size               int                                      32000
    num_key_value_heads      Optional[int]                            None
    rope_theta               float                                    10000
    rope_traditional         bool                                     False
    rope_scaling             Optional[Dict[str,Union[float,str]]]     None
    max_position_embeddings  int                                      4096
    tie_word_embeddings      bool                                     False
"""
_VISION = """
    model_type               str     'clip_vision_model'
    num_hidden_layers        int     24
    hidden_size              int     1024
    intermediate_size        int     4096
    num_attention_heads      int     16
    image_size               int     336
    patch_size               int     14
    projection_dim           int     768
    vocab_size               int     32000
    num_channels             int     3
    layer_norm_eps           float   1e-5
"""
_MODEL = """
    text_config                      object                -
    vision_config                    object                -
    model_type                       str                   'llava'
    ignore_index                     int                   -100
    image_token_index                int                   32000
    vision_feature_select_strategy   str                   'default'
    vision_feature_layer             int                   -2
    vocab_size                       int                   32000
    eos_token_id                     Optional[List[int]]   None
"""


def llama_text_rules(self):
    """MHA when kv heads are not given; only linear rope scaling, with its factor (config.py:24-37)"""
    if self.num_key_value_heads is None:
        self.num_key_value_heads = self.num_attention_heads
    scaling = self.rope_scaling
    if scaling:
        if not {"factor", "type"} <= set(scaling):
            raise ValueError("rope_scaling must contain keys {'factor', 'type'}")
        if scaling["type"] != "linear":
            raise ValueError("rope_scaling 'type' currently only supports 'linear'")


TextConfig = config_class("TextConfig", __name__, _TEXT, llama_text_rules)
VisionConfig = config_class("VisionConfig", __name__, _VISION)
ModelConfig = config_class("ModelConfig", __name__, _MODEL,
                           members={"from_dict": nested_from_dict(text_config=TextConfig, vision_config=VisionConfig)})


def llava_15_7b_config() -> "ModelConfig":
    """LLaVA-1.5-7B (CLIP-ViT-L/14-336 + Vicuna-7B; BASELINE config 3)."""
    return ModelConfig(text_config=TextConfig(vocab_size=32064, rms_norm_eps=1e-5), vision_config=VisionConfig(),
                       vocab_size=32064)
