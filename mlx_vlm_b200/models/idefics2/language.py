"""Mistral language model of Idefics2 (reference mlx_vlm/models/idefics2/language.py:16-150) on the shared
decoder engine (a Mistral layer = the Qwen2 layer with zero q/k/v bias; nn.RoPE = one rotary axis)."""
from ..llava.language import LanguageModel  # noqa: F401  (same 1-D position bookkeeping)
