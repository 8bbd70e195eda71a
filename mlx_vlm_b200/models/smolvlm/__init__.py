"""SmolVLM = Idefics3 under another model_type (reference mlx_vlm/models/smolvlm/smolvlm.py:1-6) with its own
config defaults (smolvlm/config.py:8-80)."""
from .config import ModelConfig, TextConfig, VisionConfig
from ..idefics3 import LanguageModel, VisionModel
from ..idefics3 import Model as _Idefics3Model


class Model(_Idefics3Model):
    pass
