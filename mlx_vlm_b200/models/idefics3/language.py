"""Llama language model of Idefics3 (reference mlx_vlm/models/idefics3/language.py:16-140) on the shared decoder
engine (a Llama layer = the Qwen2 layer with zero q/k/v bias; nn.RoPE = one rotary axis)."""
from ..llava.language import LanguageModel  # noqa: F401  (same 1-D position bookkeeping)
