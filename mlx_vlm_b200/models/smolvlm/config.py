"""SmolVLM configuration: the schema of reference mlx_vlm/models/smolvlm/config.py:8-80 as tables, with its derived defaults
(heads from head_dim, tower depth / MLP width from the tower width)."""
from __future__ import annotations

from ..config_schema import config_class, image_token_alias, nested_from_dict

_TEXT = """
    model_type               str             'smolvlm'
    hidden_size              int             4096
    intermediate_size        int             11008
    num_attention_heads      Optional[int]   None
    rms_norm_eps             float           1e-5
    vocab_size               int             49152
    num_key_value_heads      Optional[int]   None
    head_dim                 Optional[int]   None
    rope_theta               float           1000000.0
    num_hidden_layers        int             32
    rope_traditional         bool            False
    max_position_embeddings  int             4096
    tie_word_embeddings      bool            False
"""
_VISION = """
    model_type            str             'siglip_vision_model'
    hidden_size           Optional[int]   None
    num_attention_heads   Optional[int]   None
    patch_size            int             14
    num_hidden_layers     Optional[int]   None
    intermediate_size     Optional[int]   None
    image_size            int             384
    num_channels          int             3
    layer_norm_eps        float           1e-6
"""
_MODEL = """
    text_config         object                -
    vision_config       object                -
    model_type          str                   'smolvlm'
    ignore_index        int                   -100
    vocab_size          int                   49152
    scale_factor        int                   2
    image_token_id      int                   49153
    image_token_index   Optional[int]         None
    eos_token_id        Optional[List[int]]   None
"""
_TOWER_MLP = {768: 3072, 1152: 4304}     # SigLIP-B / SigLIP-SO400M


def _text_rules(self):
    """config.py:24-31: heads = hidden / head_dim when only head_dim is given (else 32); MHA by default"""
    if self.num_attention_heads is None:
        self.num_attention_heads = self.hidden_size // self.head_dim if (self.head_dim or 0) > 0 else 32
    if self.num_key_value_heads is None:
        self.num_key_value_heads = self.num_attention_heads


def _vision_rules(self):
    """config.py:46-66: the SmolVLM2 towers are SigLIP-B (12 layers) or SigLIP-SO400M (27 layers)"""
    if self.hidden_size is None:
        self.hidden_size = 1152
    if self.num_attention_heads is None:
        self.num_attention_heads = self.hidden_size // 64 if self.hidden_size % 64 == 0 else 16
    if self.num_hidden_layers is None:
        self.num_hidden_layers = 12 if self.hidden_size <= 768 else 27
    if self.intermediate_size is None:
        self.intermediate_size = _TOWER_MLP.get(self.hidden_size, self.hidden_size * 4)


TextConfig = config_class("TextConfig", __name__, _TEXT, _text_rules)
VisionConfig = config_class("VisionConfig", __name__, _VISION, _vision_rules)
# the reference gives `text_config` / `vision_config` default factories (config.py:71-72)
ModelConfig = config_class("ModelConfig", __name__, _MODEL, image_token_alias,
                           {"from_dict": nested_from_dict(text_config=TextConfig, vision_config=VisionConfig)},
                           factories={"text_config": TextConfig, "vision_config": VisionConfig})
