"""Qwen2.5-VL vision tower (reference mlx_vlm/models/qwen2_5_vl/vision.py:74-413).

Host integer logic (bit-exact, pinned by tests/golden/qwen2_5_vl_golden.json which executes the reference's own
source): the window permutation of the 2 x 2 merge units (`get_window_index` :258-319), the de-duplicated window
boundaries and the per-frame boundaries (:331-366), the (row, column) rotary ids (:226-256).

Device: patch-embed GEMM, window gather, depth x {RMSNorm, qkv + bias, 2-D rotary, attention inside ragged segments
(windows, or whole frames in `fullatt_block_indexes`), proj, RMSNorm, SwiGLU MLP with biases}, merger (RMSNorm, Linear,
erf-GELU, Linear), reverse gather — split-operand tensor-core GEMMs and fp32 row kernels (models/tower_ops.py).  The
reference casts the pixels to the weight dtype and runs this tower in bf16; here it is computed at fp32 accuracy from
the bf16-rounded pixels (a superset: the result differs from the reference's by the reference's own bf16 rounding noise,
which is the tolerance of the parity test) and rounded to bf16 once, where the features are merged."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from ... import _native as N
from ..qwen2_vl.vision import check_array_shape  # noqa: F401  (same layout rule, vision.py:10-27)
from ..tower_ops import EPI_GELU_EXACT, SplitBuf, TowerOps
from .config import VisionConfig


def rot_pos_ids(grid_thw, merge: int) -> np.ndarray:
    """vision.py:226-250: (N, 2) (row, column) per patch, patches ordered merge unit by merge unit"""
    out = []
    for t, h, w in np.asarray(grid_thw).tolist():
        rows = np.repeat(np.arange(h)[:, None], w, axis=1)
        cols = np.repeat(np.arange(w)[None, :], h, axis=0)
        rows = rows.reshape(h // merge, merge, w // merge, merge).transpose(0, 2, 1, 3).reshape(-1)
        cols = cols.reshape(h // merge, merge, w // merge, merge).transpose(0, 2, 1, 3).reshape(-1)
        out.append(np.tile(np.stack([rows, cols], -1), (t, 1)))
    return np.concatenate(out, 0)


def get_window_index(grid_thw, window_size: int, patch_size: int, merge: int) -> Tuple[np.ndarray, np.ndarray]:
    """vision.py:258-319 -> (merge-unit order after grouping by window, window boundaries in patches incl. empty windows)"""
    side = window_size // merge // patch_size
    order: List[np.ndarray] = []
    bounds = [0]
    first = 0
    for t, h, w in np.asarray(grid_thw).tolist():
        uh, uw = h // merge, w // merge
        units = np.arange(t * uh * uw).reshape(t, uh, uw)
        extra_h, extra_w = side - uh % side, side - uw % side      # the reference adds a whole empty window when it divides
        wh, ww = (uh + extra_h) // side, (uw + extra_w) // side
        padded = np.pad(units, ((0, 0), (0, extra_h), (0, extra_w)), constant_values=-100)
        padded = padded.reshape(t, wh, side, ww, side).transpose(0, 1, 3, 2, 4).reshape(t, wh * ww, side * side)
        real = padded != -100
        order.append(padded[real] + first)
        bounds.extend((np.cumsum(real.sum(-1).reshape(-1)) * merge * merge + bounds[-1]).tolist())
        first += t * uh * uw
    return np.concatenate(order), np.asarray(bounds, dtype=np.int64)


def segment_tables(grid_thw, cfg: VisionConfig):
    """-> (window order, window boundaries without repeats, frame boundaries)  (vision.py:331-366)"""
    order, raw = get_window_index(grid_thw, cfg.window_size, cfg.patch_size, cfg.spatial_merge_size)
    _, first = np.unique(raw, return_index=True)
    windows = raw[np.sort(first)]
    frames = [0]
    for t, h, w in np.asarray(grid_thw).tolist():
        for _ in range(t):
            frames.append(frames[-1] + h * w)
    return order, windows, np.asarray(frames, dtype=np.int64)


class VisionModel:
    def __init__(self, config: VisionConfig, engine_getter):
        self.config = config
        self.model_type = config.model_type
        if self.model_type != "qwen2_5_vl":
            raise ValueError(f"Unsupported model type: {self.model_type}")
        self.spatial_merge_size = config.spatial_merge_size
        self._engine = engine_getter
        self.w: Dict[str, torch.Tensor] = {}

    def sanitize(self, weights):
        out = {}
        for k, v in weights.items():
            if "position_ids" in k:
                continue
            if "patch_embed.proj.weight" in k and v.ndim == 5 and v.shape[-1] == self.config.in_channels \
                    and v.shape[1] != self.config.in_channels:
                v = v.permute(0, 4, 1, 2, 3)      # MLX [O,T,H,W,C] -> [O,C,T,H,W]: the pixel rows are (C,T,H,W)-ordered
            out[k] = v
        return out

    def load(self, weights: Dict[str, torch.Tensor], prefix: str = "vision_tower."):
        c, eng = self.config, self._engine()
        dev = eng.device

        def put(name, t):
            t = t.to(device=dev, dtype=torch.bfloat16)
            if t.ndim == 2 and t.shape[1] % 8:      # TMA needs a 16-byte row pitch (SwiGLU width 3420 -> 3424 columns)
                t = torch.nn.functional.pad(t, (0, 8 - t.shape[1] % 8))
            self.w[name] = t.contiguous()

        E = c.hidden_size
        put("patch", weights[prefix + "patch_embed.proj.weight"].reshape(E, -1))
        for i in range(c.depth):
            q = prefix + f"blocks.{i}."
            put(f"{i}.n1", weights[q + "norm1.weight"]); put(f"{i}.n2", weights[q + "norm2.weight"])
            put(f"{i}.qkv.w", weights[q + "attn.qkv.weight"]); put(f"{i}.qkv.b", weights[q + "attn.qkv.bias"])
            put(f"{i}.proj.w", weights[q + "attn.proj.weight"]); put(f"{i}.proj.b", weights[q + "attn.proj.bias"])
            put(f"{i}.gu.w", torch.cat([weights[q + "mlp.gate_proj.weight"], weights[q + "mlp.up_proj.weight"]], 0))
            put(f"{i}.gu.b", torch.cat([weights[q + "mlp.gate_proj.bias"], weights[q + "mlp.up_proj.bias"]], 0))
            put(f"{i}.down.w", weights[q + "mlp.down_proj.weight"]); put(f"{i}.down.b", weights[q + "mlp.down_proj.bias"])
        m = prefix + "merger."
        put("m.ln", weights[m + "ln_q.weight"])
        put("m.fc1.w", weights[m + "mlp.0.weight"]); put("m.fc1.b", weights[m + "mlp.0.bias"])
        put("m.fc2.w", weights[m + "mlp.2.weight"]); put("m.fc2.b", weights[m + "mlp.2.bias"])
        hd = E // c.num_heads
        dim = hd // 2
        inv = (1.0 / (10000.0 ** (np.arange(0, dim, 2, dtype=np.float32) / np.float32(dim)))).astype(np.float32)
        with torch.cuda.stream(eng.stream):
            self.w["inv_freq"] = torch.from_numpy(inv).to(dev)

    def __call__(self, hidden_states: torch.Tensor, grid_thw, output_hidden_states=None):
        """pixel rows (N, C*T*ps*ps) -> merged features (N / merge^2, out_hidden) bf16, in the original patch order"""
        if output_hidden_states:
            raise NotImplementedError("output_hidden_states is not produced by this tower")
        c, eng = self.config, self._engine()
        ops = TowerOps(eng)
        grid = grid_thw.cpu().numpy() if isinstance(grid_thw, torch.Tensor) else np.asarray(grid_thw)
        grid = grid.reshape(-1, 3).astype(np.int64)
        E, I, nh = c.hidden_size, c.intermediate_size, c.num_heads
        hd = E // nh
        unit = c.spatial_merge_size ** 2
        T = int((grid[:, 0] * grid[:, 1] * grid[:, 2]).sum())
        w = self.w
        order, windows, frames = segment_tables(grid, c)
        pos = rot_pos_ids(grid, c.spatial_merge_size).reshape(T // unit, unit, 2)[order].reshape(T, 2)
        host = [np.ascontiguousarray(order.astype(np.int32)), np.ascontiguousarray(np.argsort(order, kind="stable").astype(np.int32)),
                np.ascontiguousarray(pos.astype(np.int32)), np.ascontiguousarray(windows.astype(np.int32)),
                np.ascontiguousarray(frames.astype(np.int32))]
        with torch.cuda.stream(eng.stream):
            x = hidden_states.to(device=eng.device, dtype=torch.float32)
            if x.shape[0] != T:
                raise ValueError(f"pixel_values has {x.shape[0]} patch rows, image_grid_thw describes {T}")
            x = x.contiguous()
            order_d, back_d, pos_d, win_d, frm_d = [torch.from_numpy(a).to(eng.device) for a in host]
        K = x.shape[1]
        xw = ops.f32(T, K)
        ops.gather_rows(x, order_d, unit, xw)                      # merge units grouped by window
        pat = SplitBuf(eng, T, K)
        ops.split(xw, pat)
        h = ops.f32(T, E)
        ops.linear(pat, w["patch"], None, out32=h, k_w=K, n_parts=1)   # pixels rounded to bf16 (qwen2_5_vl.py: astype)
        y, o, act = SplitBuf(eng, T, E), SplitBuf(eng, T, E), SplitBuf(eng, T, I)
        qkv, gu = ops.f32(T, 3 * E), ops.f32(T, 2 * I)
        win_max, frm_max = int(np.diff(windows).max()), int(np.diff(frames).max())
        for i in range(c.depth):
            full = i in c.fullatt_block_indexes
            ops.rms_norm(h, w[f"{i}.n1"], 1e-6, out_split=y)
            ops.linear(y, w[f"{i}.qkv.w"], w[f"{i}.qkv.b"], out32=qkv)
            ops.vision_rope(qkv, pos_d, w["inv_freq"], nh, hd)
            ops.attention_varlen((qkv, 3 * E, hd), (qkv[:, E:], 3 * E, hd), (qkv[:, 2 * E:], 3 * E, hd), n_heads=nh,
                                 n_kv=nh, hd=hd, cu=frm_d if full else win_d,
                                 n_seg=(len(frames) if full else len(windows)) - 1,
                                 max_len=frm_max if full else win_max, scale=hd ** -0.5, out_split=o)
            ops.linear(o, w[f"{i}.proj.w"], w[f"{i}.proj.b"], out32=h, res32=h)
            ops.rms_norm(h, w[f"{i}.n2"], 1e-6, out_split=y)
            ops.linear(y, w[f"{i}.gu.w"], w[f"{i}.gu.b"], out32=gu)
            ops.swiglu(gu, act)
            ops.linear(act, w[f"{i}.down.w"], w[f"{i}.down.b"], out32=h, res32=h)
        Tm, Em = T // unit, E * unit
        n32 = ops.f32(T, E)
        ops.rms_norm(h, w["m.ln"], 1e-6, out32=n32)
        ms, mid = SplitBuf(eng, Tm, Em), SplitBuf(eng, Tm, Em)
        ops.split(n32.view(Tm, Em), ms)
        ops.linear(ms, w["m.fc1.w"], w["m.fc1.b"], out_split=mid, epi=EPI_GELU_EXACT)
        merged = ops.f32(Tm, w["m.fc2.w"].shape[0])
        ops.linear(mid, w["m.fc2.w"], w["m.fc2.b"], out32=merged)
        unwound = ops.f32(*merged.shape)
        ops.gather_rows(merged, back_d, 1, unwound)                # back to the order of the image tokens
        feats = eng.empty(tuple(unwound.shape))
        N.check(eng.lib.b200_cast_f32_bf16(unwound.data_ptr(), feats.data_ptr(), unwound.numel(), eng.s), "cast")
        return feats
