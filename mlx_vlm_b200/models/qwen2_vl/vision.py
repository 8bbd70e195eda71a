"""Qwen2-VL vision tower — Python face of the reference's
mlx_vlm/models/qwen2_vl/vision.py (`VisionModel.__call__` :257-290, `sanitize`
:292-310).  The arithmetic (patch-embed GEMM, 32 x {LN, QKV, 2-D rotary,
attention, proj, LN, MLP}, PatchMerger) runs in libb200vlm.so
(`b200_engine_vision`)."""
from __future__ import annotations

import numpy as np
import torch

from .config import VisionConfig


def check_array_shape(arr):
    """vision.py:9-25: is the conv weight already in the MLX [O, T, H, W, C] layout?"""
    shape = arr.shape
    if len(shape) not in [4, 5]:
        return False
    B, out_channels, kH, KW, t = shape
    if t == 3:
        return True
    return (out_channels >= kH) and (out_channels >= KW) and (kH == KW)


class VisionModel:
    def __init__(self, config: VisionConfig, engine_getter):
        self.config = config
        self.model_type = config.model_type
        if self.model_type != "qwen2_vl":
            raise ValueError(f"Unsupported model type: {self.model_type}")
        self.spatial_merge_size = config.spatial_merge_size
        self._engine = engine_getter

    def __call__(self, hidden_states: torch.Tensor, grid_thw, output_hidden_states=None):
        if output_hidden_states:
            raise NotImplementedError("output_hidden_states is not produced by the fused tower")
        grid = grid_thw.cpu().numpy() if isinstance(grid_thw, torch.Tensor) else np.asarray(grid_thw)
        eng = self._engine()
        if hidden_states.dtype != torch.float32 or not hidden_states.is_cuda:
            hidden_states = hidden_states.to(device=eng.device, dtype=torch.float32)
        return eng.vision(hidden_states, grid)

    def sanitize(self, weights):
        out = {}
        for k, v in weights.items():
            if "position_ids" in k:
                continue
            elif "patch_embed.proj.weight" in k:
                # keep the HF [O, C, T, H, W] layout: the patch rows are ordered (C,T,H,W)
                if v.ndim == 5 and v.shape[-1] == 3 and v.shape[1] != 3:
                    v = v.permute(0, 4, 1, 2, 3)  # MLX [O,T,H,W,C] -> [O,C,T,H,W]
                out[k] = v
            else:
                out[k] = v
        return out
