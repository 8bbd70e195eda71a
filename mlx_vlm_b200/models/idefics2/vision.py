"""SigLIP vision tower of Idefics2 (reference mlx_vlm/models/idefics2/vision.py:20-215), in fp32 with
bf16-valued weights like the reference runs it (float32 `pixel_values` are never cast, idefics2.py:212-251).

Host integer logic (bit-exact, pinned by tests/golden): the bucketed fractional position ids
(vision.py:150-173).  Device: patch embedding with bias, position gather, N x {LN, MHA, LN, fast-GELU MLP},
post-LayerNorm — split-operand tensor-core GEMMs and fp32 kernels (models/tower_ops.py)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from ..tower_ops import EPI_GELU_FAST, SplitBuf, TowerOps
from .config import VisionConfig


def bucketed_position_ids(patch_mask: np.ndarray, side: int) -> np.ndarray:
    """vision.py:150-170.  patch_mask (B, ph, pw) bool -> (B, ph * pw) int64.  The valid nh x nw block of
    an image maps onto the side x side position grid through its fractional coordinates; the reference
    uses `np.digitize(frac, boundaries, right=True) - 1`, so coordinate 0 falls into bucket -1 and those
    ids are negative (they index the position table from its end, like the reference's mx gather)."""
    m = np.asarray(patch_mask).astype(bool)
    B, ph, pw = m.shape
    edges = np.linspace(1 / side, 1.0, side, endpoint=False)
    ids = np.zeros((B, ph * pw), dtype=np.int64)
    for b in range(B):
        rows, cols = int(m[b, :, 0].sum()), int(m[b, 0, :].sum())
        r = np.digitize(np.linspace(0, 1, rows, endpoint=False), edges, right=True) - 1
        c = np.digitize(np.linspace(0, 1, cols, endpoint=False), edges, right=True) - 1
        ids[b, m[b].reshape(-1)] = (r[:, None] * side + c[None, :]).reshape(-1)
    return ids


class VisionModel:
    def __init__(self, config: VisionConfig, engine_getter):
        self.config = config
        self.model_type = config.model_type
        self._engine = engine_getter
        self.w: Dict[str, torch.Tensor] = {}

    def sanitize(self, weights):
        out = {}
        for k, v in weights.items():
            if "position_ids" in k:
                continue
            if "patch_embedding.weight" in k and v.ndim == 4 and v.shape[1] == self.config.num_channels \
                    and v.shape[-1] != self.config.num_channels:
                v = v.permute(0, 2, 3, 1)     # PyTorch [O, C, kH, kW] -> [O, kH, kW, C]
            out[k] = v
        return out

    def load(self, weights: Dict[str, torch.Tensor], prefix: str = "vision_model."):
        c, eng = self.config, self._engine()
        E, dev = c.hidden_size, eng.device

        def put(name, t):
            self.w[name] = t.to(device=dev, dtype=torch.bfloat16).contiguous()

        K = c.patch_size * c.patch_size * c.num_channels
        conv = weights[prefix + "embeddings.patch_embedding.weight"].reshape(E, K)
        convp = torch.zeros(E, (K + 7) // 8 * 8, dtype=conv.dtype)
        convp[:, :K] = conv
        put("patch", convp)
        put("patch.b", weights[prefix + "embeddings.patch_embedding.bias"])
        put("pos", weights[prefix + "embeddings.position_embedding.weight"])
        put("post.w", weights[prefix + "post_layernorm.weight"]); put("post.b", weights[prefix + "post_layernorm.bias"])
        for i in range(c.num_hidden_layers):
            q = prefix + f"encoder.layers.{i}."
            put(f"{i}.ln1.w", weights[q + "layer_norm1.weight"]); put(f"{i}.ln1.b", weights[q + "layer_norm1.bias"])
            put(f"{i}.ln2.w", weights[q + "layer_norm2.weight"]); put(f"{i}.ln2.b", weights[q + "layer_norm2.bias"])
            put(f"{i}.qkv.w", torch.cat([weights[q + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0))
            put(f"{i}.qkv.b", torch.cat([weights[q + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0))
            put(f"{i}.out.w", weights[q + "self_attn.out_proj.weight"]); put(f"{i}.out.b", weights[q + "self_attn.out_proj.bias"])
            put(f"{i}.fc1.w", weights[q + "mlp.fc1.weight"]); put(f"{i}.fc1.b", weights[q + "mlp.fc1.bias"])
            put(f"{i}.fc2.w", weights[q + "mlp.fc2.weight"]); put(f"{i}.fc2.b", weights[q + "mlp.fc2.bias"])

    def __call__(self, x: torch.Tensor, patch_attention_mask: Optional[np.ndarray] = None,
                 output_hidden_states: Optional[bool] = None):
        """x: NHWC fp32 (n_img, H, W, C) on the device -> (pooler_output fp32 (n_img * P, E), None, None).
        The reference calls the encoder WITHOUT an attention mask (vision.py:207): padded patches take
        part in attention, only their position id stays 0."""
        c, eng = self.config, self._engine()
        ops = TowerOps(eng)
        B, H, W, C = x.shape
        ps, E, I = c.patch_size, c.hidden_size, c.intermediate_size
        gh, gw = H // ps, W // ps
        P = gh * gw
        T = B * P
        nh = c.num_attention_heads
        hd = E // nh
        K = ps * ps * C
        w = self.w
        if patch_attention_mask is None:
            patch_attention_mask = np.ones((B, gh, gw), dtype=bool)
        pos_ids = bucketed_position_ids(patch_attention_mask, c.image_size // ps).astype(np.int32)
        pos_dev = torch.from_numpy(np.ascontiguousarray(pos_ids))
        with torch.cuda.stream(eng.stream):
            pos_dev = pos_dev.to(eng.device)
        x = x.contiguous()
        pat = SplitBuf(eng, T, K)
        ops.patchify(x, ps, pat)
        patch = ops.f32(T, E)
        ops.linear(pat, w["patch"], w["patch.b"], out32=patch, k_w=K)
        h = ops.f32(T, E)
        ops.embed(patch, None, w["pos"], pos_dev, h, B, P)
        y, o, mlp = SplitBuf(eng, T, E), SplitBuf(eng, T, E), SplitBuf(eng, T, I)
        qkv = ops.f32(T, 3 * E)
        for i in range(c.num_hidden_layers):
            ops.layer_norm(h, w[f"{i}.ln1.w"], w[f"{i}.ln1.b"], c.layer_norm_eps, out_split=y)
            ops.linear(y, w[f"{i}.qkv.w"], w[f"{i}.qkv.b"], out32=qkv)
            ops.attention((qkv, 3 * E, hd), (qkv[:, E:], 3 * E, hd), (qkv[:, 2 * E:], 3 * E, hd), n_heads=nh, n_kv=nh,
                          hd=hd, Lq=P, S=P, n_seg=B, q_seg=P, k_seg=P, scale=hd ** -0.5, out_split=o)
            ops.linear(o, w[f"{i}.out.w"], w[f"{i}.out.b"], out32=h, res32=h)
            ops.layer_norm(h, w[f"{i}.ln2.w"], w[f"{i}.ln2.b"], c.layer_norm_eps, out_split=y)
            ops.linear(y, w[f"{i}.fc1.w"], w[f"{i}.fc1.b"], out_split=mlp, epi=EPI_GELU_FAST)
            ops.linear(mlp, w[f"{i}.fc2.w"], w[f"{i}.fc2.b"], out32=h, res32=h)
        pooled = ops.f32(T, E)
        ops.layer_norm(h, w["post.w"], w["post.b"], 1e-5, out32=pooled)   # nn.LayerNorm default eps (vision.py:184)
        return pooled, None, None
