"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's Qwen2.5-VL generate path.  The language model, the M-RoPE index, the
merge and the generation loop are Qwen2-VL's (models/qwen2_5_vl/language.py and qwen2_5_vl.py differ from qwen2_vl's
only in comments and in the tiling of text-only position ids) and are taken from oracle/qwen2vl.py; the vision tower is
restated here:
  models/qwen2_5_vl/vision.py:74-103   PatchEmbed (Conv3d kernel == stride, no bias)
  models/qwen2_5_vl/vision.py:106-121  PatchMerger: RMSNorm(1e-6) -> Linear -> GELU (erf) -> Linear
  models/qwen2_5_vl/vision.py:124-166  Attention: qkv with bias, 2-D rotary (fp32 cos / sin, one cast), SDPA per
                                       segment of `cu_seqlens`, proj
  models/qwen2_5_vl/vision.py:169-195  MLP (SwiGLU, all three Linears with bias), block = RMSNorm(1e-6) pre-norm residual
  models/qwen2_5_vl/vision.py:226-256  rot_pos_emb
  models/qwen2_5_vl/vision.py:258-319  get_window_index
  models/qwen2_5_vl/vision.py:321-389  __call__: window permutation of the merge units, windowed attention except in
                                       `fullatt_block_indexes`, merger, reverse permutation
The integer logic is pinned by executing the reference's own source (tests/golden/make_qwen2_5_vl_golden.py); the
floating-point rounding points are those of oracle/mlx_semantics.py (unpinned at the mlx boundary, stated there)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder


@dataclass
class VisionCfg:
    depth: int = 32
    hidden_size: int = 1280
    intermediate_size: int = 3420
    out_hidden_size: int = 1536
    num_heads: int = 16
    patch_size: int = 14
    in_channels: int = 3
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: Tuple[int, ...] = (7, 15, 23, 31)


@dataclass
class Cfg:
    text: Q.TextCfg = field(default_factory=lambda: Q.TextCfg(hidden_size=2048, num_hidden_layers=36, intermediate_size=11008,
                                                              num_attention_heads=16, num_key_value_heads=2))
    vision: VisionCfg = field(default_factory=lambda: VisionCfg(out_hidden_size=2048))
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653


def tiny_cfg() -> Cfg:
    t = Q.TextCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                  num_key_value_heads=2, vocab_size=1024, mrope_section=(8, 12, 12), tie_word_embeddings=True)
    v = VisionCfg(depth=3, hidden_size=128, intermediate_size=192, out_hidden_size=256, num_heads=4, window_size=56,
                  fullatt_block_indexes=(1,))
    return Cfg(text=t, vision=v, image_token_id=1000, video_token_id=1001, vision_start_token_id=1002,
               vision_end_token_id=1003)


def _qcfg(cfg: Cfg) -> Q.Cfg:
    """the Qwen2-VL view of the text side (same ids, same LM)"""
    qv = Q.VisionCfg(depth=cfg.vision.depth, embed_dim=cfg.vision.hidden_size, hidden_size=cfg.vision.out_hidden_size,
                     num_heads=cfg.vision.num_heads, patch_size=cfg.vision.patch_size, in_channels=cfg.vision.in_channels,
                     spatial_merge_size=cfg.vision.spatial_merge_size, temporal_patch_size=cfg.vision.temporal_patch_size)
    return Q.Cfg(text=cfg.text, vision=qv, image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id,
                 vision_start_token_id=cfg.vision_start_token_id, vision_end_token_id=cfg.vision_end_token_id)


# ---------------------------------------------------------------------------------------------- integer logic
def get_window_index(grid_thw, window_size: int, patch_size: int, merge: int):
    """vision.py:258-319 -> (window_index over merge units, raw cu_window_seqlens in patches)"""
    win = window_size // merge // patch_size
    index_all: List[np.ndarray] = []
    cu = [0]
    base = 0
    for t, h, w in np.asarray(grid_thw).tolist():
        gh, gw = h // merge, w // merge
        idx = np.arange(t * gh * gw).reshape(t, gh, gw)
        ph, pw = win - gh % win, win - gw % win          # a full extra window when the grid divides evenly (reference)
        nh, nw = (gh + ph) // win, (gw + pw) // win
        pad = np.pad(idx, ((0, 0), (0, ph), (0, pw)), constant_values=-100)
        pad = pad.reshape(t, nh, win, nw, win).transpose(0, 1, 3, 2, 4).reshape(t, nh * nw, win, win)
        lens = (pad != -100).sum(axis=(2, 3)).reshape(-1)
        flat = pad.reshape(-1)
        index_all.append(flat[flat != -100] + base)
        cu.extend((np.cumsum(lens) * merge * merge + cu[-1]).tolist())
        base += t * gh * gw
    return np.concatenate(index_all), np.asarray(cu, dtype=np.int64)


def segment_tables(grid_thw, vcfg: VisionCfg):
    """-> window_index, de-duplicated window boundaries, per-frame boundaries (vision.py:331-366)"""
    widx, cu_raw = get_window_index(grid_thw, vcfg.window_size, vcfg.patch_size, vcfg.spatial_merge_size)
    keep, seen = [], set()
    for i, x in enumerate(cu_raw.tolist()):
        if x not in seen:
            seen.add(x)
            keep.append(i)
    cu_win = cu_raw[keep]
    frames = []
    for t, h, w in np.asarray(grid_thw).tolist():
        frames += [h * w] * t
    cu_full = np.concatenate([[0], np.cumsum(frames)]).astype(np.int64)
    return widx, cu_win, cu_full


def vision_rotary_freqs(grid_thw, vcfg: VisionCfg) -> torch.Tensor:
    qv = Q.VisionCfg(embed_dim=vcfg.hidden_size, num_heads=vcfg.num_heads, spatial_merge_size=vcfg.spatial_merge_size)
    return Q.vision_rotary_freqs(grid_thw, qv)


# ---------------------------------------------------------------------------------------------- weights
def weight_shapes(cfg: Cfg) -> Dict[str, Tuple[int, ...]]:
    v = cfg.vision
    E, I = v.hidden_size, v.intermediate_size
    s: Dict[str, Tuple[int, ...]] = {}
    s["vision_tower.patch_embed.proj.weight"] = (E, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
    for i in range(v.depth):
        p = f"vision_tower.blocks.{i}."
        s[p + "norm1.weight"], s[p + "norm2.weight"] = (E,), (E,)
        s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"] = (3 * E, E), (3 * E,)
        s[p + "attn.proj.weight"], s[p + "attn.proj.bias"] = (E, E), (E,)
        s[p + "mlp.gate_proj.weight"], s[p + "mlp.gate_proj.bias"] = (I, E), (I,)
        s[p + "mlp.up_proj.weight"], s[p + "mlp.up_proj.bias"] = (I, E), (I,)
        s[p + "mlp.down_proj.weight"], s[p + "mlp.down_proj.bias"] = (E, I), (E,)
    m = v.spatial_merge_size ** 2 * E
    s["vision_tower.merger.ln_q.weight"] = (E,)
    s["vision_tower.merger.mlp.0.weight"], s["vision_tower.merger.mlp.0.bias"] = (m, m), (m,)
    s["vision_tower.merger.mlp.2.weight"], s["vision_tower.merger.mlp.2.bias"] = (v.out_hidden_size, m), (v.out_hidden_size,)
    for k, shp in Q.weight_shapes(_qcfg(cfg)).items():
        if k.startswith("language_model."):
            s[k] = shp
    return s


def init_weights(cfg: Cfg, seed: int = 0, std: float = 0.02, norm_jitter: float = 0.05) -> Dict[str, torch.Tensor]:
    R = Rounder("bf16")
    out = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        if ("norm" in name or "ln_q" in name) and len(shape) == 1:
            w = torch.ones(shape) + norm_jitter * torch.randn(shape, generator=g)
        else:
            w = std * torch.randn(shape, generator=g)
        out[name] = R.r(w)
    return out


# ---------------------------------------------------------------------------------------------- tower
def vision_forward(cfg: Cfg, W: Dict[str, torch.Tensor], pixel_values, grid_thw, R: Rounder):
    """pixel_values (N, C*tps*ps*ps) f32 -> merged features (N/4, out_hidden) in the ORIGINAL merge-unit order"""
    v = cfg.vision
    E, nh = v.hidden_size, v.num_heads
    hd = E // nh
    unit = v.spatial_merge_size ** 2
    x = R.r(torch.as_tensor(pixel_values, dtype=torch.float32))       # qwen2_5_vl.py: astype(weight dtype)
    h = S.linear(R, x, W["vision_tower.patch_embed.proj.weight"].reshape(E, -1))
    freqs = vision_rotary_freqs(grid_thw, v)
    widx, cu_win, cu_full = segment_tables(grid_thw, v)
    N = h.shape[0]
    perm = torch.from_numpy(widx)
    h = h.reshape(N // unit, unit, E)[perm].reshape(N, E)
    freqs = freqs.reshape(N // unit, unit, -1)[perm].reshape(N, -1)
    cos = torch.cos(freqs).repeat(1, 2)[:, None, :]
    sin = torch.sin(freqs).repeat(1, 2)[:, None, :]
    scale = hd ** -0.5
    for i in range(v.depth):
        p = f"vision_tower.blocks.{i}."
        cu = cu_full if i in v.fullatt_block_indexes else cu_win
        y = S.rms_norm(R, h, W[p + "norm1.weight"], 1e-6)
        qkv = S.linear(R, y, W[p + "attn.qkv.weight"], W[p + "attn.qkv.bias"]).reshape(N, 3, nh, hd)
        q, k, vv = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q = R.r(q * cos + Q._rotate_half(q) * sin)
        k = R.r(k * cos + Q._rotate_half(k) * sin)
        outs = []
        for s in range(len(cu) - 1):
            a, b = int(cu[s]), int(cu[s + 1])
            o = S.sdpa(R, q[a:b].transpose(0, 1)[None], k[a:b].transpose(0, 1)[None], vv[a:b].transpose(0, 1)[None],
                       scale, causal=False)
            outs.append(o[0].transpose(0, 1).reshape(b - a, E))
        att = S.linear(R, torch.cat(outs, 0), W[p + "attn.proj.weight"], W[p + "attn.proj.bias"])
        h = R.r(h + att)
        y = S.rms_norm(R, h, W[p + "norm2.weight"], 1e-6)
        g = S.silu(R, S.linear(R, y, W[p + "mlp.gate_proj.weight"], W[p + "mlp.gate_proj.bias"]))
        u = S.linear(R, y, W[p + "mlp.up_proj.weight"], W[p + "mlp.up_proj.bias"])
        y = S.linear(R, R.r(g * u), W[p + "mlp.down_proj.weight"], W[p + "mlp.down_proj.bias"])
        h = R.r(h + y)
    y = S.rms_norm(R, h, W["vision_tower.merger.ln_q.weight"], 1e-6).reshape(-1, E * unit)
    y = S.gelu_exact(R, S.linear(R, y, W["vision_tower.merger.mlp.0.weight"], W["vision_tower.merger.mlp.0.bias"]))
    y = S.linear(R, y, W["vision_tower.merger.mlp.2.weight"], W["vision_tower.merger.mlp.2.bias"])
    return y[torch.from_numpy(np.argsort(widx, kind="stable"))]


def greedy_generate(cfg: Cfg, W, input_ids, pixel_values, grid_thw, max_tokens: int, dtype: str = "bf16"):
    """generate_step (ar.py:151-515), greedy, EOS ignored -> dict(tokens, logits per step, image_features, inputs_embeds)"""
    R = Rounder(dtype)
    qc = _qcfg(cfg)
    ids = np.asarray(input_ids, dtype=np.int64)
    B, T = ids.shape
    emb = W["language_model.model.embed_tokens.weight"][torch.as_tensor(ids)]
    feats = None
    if pixel_values is None:
        pos, deltas = Q.get_rope_index(qc, ids)
    else:
        feats = vision_forward(cfg, W, pixel_values, grid_thw, R)
        emb = Q.merge_input_ids_with_image_features(qc, feats, emb, ids)
        pos, deltas = Q.get_rope_index(qc, ids, grid_thw, None, None)
    cache = [Q.OracleKVCache() for _ in range(cfg.text.num_hidden_layers)]
    hidden = Q.lm_layers_forward(qc, W, emb, pos, cache, R)
    logits = Q.lm_head(qc, W, hidden[:, -1, :], R)
    toks, all_logits = [], []
    for n in range(max_tokens):
        y = S.argmax_lowest(Q.logprobs_from_logits(R, logits))
        toks.append(int(y[0]))
        all_logits.append(logits.clone())
        if n == max_tokens - 1:
            break
        e = W["language_model.model.embed_tokens.weight"][y][:, None, :]
        hidden = Q.lm_layers_forward(qc, W, e, Q.decode_position_ids(cache[0].offset, deltas, B), cache, R)
        logits = Q.lm_head(qc, W, hidden[:, -1, :], R)
    return {"tokens": toks, "logits": all_logits, "image_features": feats, "inputs_embeds": emb}


def synthetic_request(cfg: Cfg, n_text: int, grid_hw=(8, 12), seed: int = 0):
    """random normalised patches for one image of grid (1, h, w) + a prompt with its placeholder expanded"""
    rng = np.random.default_rng(seed)
    v = cfg.vision
    h, w = grid_hw
    pv = rng.standard_normal((h * w, v.in_channels * v.temporal_patch_size * v.patch_size ** 2)).astype(np.float32)
    grid = np.asarray([[1, h, w]], dtype=np.int64)
    text = rng.integers(0, min(cfg.text.vocab_size, cfg.image_token_id) - 16, size=n_text).tolist()
    n_img = h * w // v.spatial_merge_size ** 2
    ids = text[:3] + [cfg.vision_start_token_id] + [cfg.image_token_id] * n_img + [cfg.vision_end_token_id] + text[3:]
    return dict(pixel_values=pv, image_grid_thw=grid, input_ids=np.asarray([ids], dtype=np.int64))
