"""CLIP tower of LLaVA-Next (reference mlx_vlm/models/llava_next/vision.py): identical to LLaVA-1.5's, run in fp32 with
bf16-valued weights on the crops of the image."""
from ..llava.vision import VisionModel  # noqa: F401
