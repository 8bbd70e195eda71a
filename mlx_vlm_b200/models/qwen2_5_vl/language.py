"""Qwen2.5-VL language model: Qwen2-VL's (reference mlx_vlm/models/qwen2_5_vl/language.py differs from
qwen2_vl/language.py only in comments and in the batch tiling of text-only position ids) on the shared decoder engine."""
from ..qwen2_vl.language import LanguageModel  # noqa: F401
