"""LLaVA-Next configuration: the schema of reference mlx_vlm/models/llava_next/config.py:8-60 as tables (the Mistral-7B
language model: 8 kv heads, rope_theta 1e6, intermediate 14336, vocabulary 32064; the CLIP tower of LLaVA-1.5)."""
from __future__ import annotations

from ..config_schema import config_class, nested_from_dict
from ..llava.config import VisionConfig, llama_text_rules  # noqa: F401

_TEXT = """
    model_type               str                                      'mistral'
    hidden_size              int                                      4096
    num_hidden_layers        int                                      32
    intermediate_size        int                                      14336
    num_attention_heads      int                                      32
    rms_norm_eps             float                                    1e-05
    vocab_size               int                                      32064
    num_key_value_heads      Optional[int]                            8
    rope_theta               float                                    1000000
    rope_traditional         bool                                     False
    rope_scaling             Optional[Dict[str,Union[float,str]]]     None
    max_position_embeddings  int                                      4096
"""
_MODEL = """
    text_config                      object                -
    vision_config                    object                -
    model_type                       str                   'llava_next'
    ignore_index                     int                   -100
    image_token_index                int                   32000
    vision_feature_select_strategy   str                   'default'
    vision_feature_layer             int                   -2
    vocab_size                       int                   32000
    eos_token_id                     Optional[List[int]]   None
"""

# the reference's LLaVA-Next text config has no `tie_word_embeddings` (the head is never tied): a constant for the shared code
TextConfig = config_class("TextConfig", __name__, _TEXT, llama_text_rules, {"tie_word_embeddings": False})
ModelConfig = config_class("ModelConfig", __name__, _MODEL,
                           members={"from_dict": nested_from_dict(text_config=TextConfig, vision_config=VisionConfig)})
