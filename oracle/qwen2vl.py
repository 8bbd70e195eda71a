"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement (torch-CPU fp32 arithmetic + explicit activation-dtype rounding,
see oracle/mlx_semantics.py) of the reference's Qwen2-VL generate path:

  reference file (under /root/reference/mlx_vlm)          restated here as
  ----------------------------------------------------    -----------------------
  models/qwen3_vl/processing_qwen3_vl.py:182-205,302-354  smart_resize, preprocess_image
  models/qwen2_vl/processing_qwen2_vl.py:93-105           expand_image_tokens
  models/qwen2_vl/vision.py:35-50,68-102,105-290          vision_forward (+rot_pos_emb)
  models/qwen2_vl/qwen2_vl.py:78-148                      merge_input_ids_with_image_features
  models/qwen2_vl/language.py:216-402                     get_rope_index
  models/qwen2_vl/language.py:40-200,404-518              lm_forward (+ decode positions)
  models/rope_utils.py:519-532,1042-1044,1227-1241,       mrope_cos_sin, apply_mrope
      1289-1334,1456-1504
  models/cache.py:337-439                                 OracleKVCache
  generate/ar.py:334-389,474-515                          greedy_generate (_step + loop)
  sample_utils.py:63-64                                   greedy sampler (argmax)

PARITY STATUS: integer functions (get_rope_index, merge indexing, decode position
bookkeeping) are pinned against the reference's own known-answer tests
(tests/test_models.py:11866-11930, tests/test_rope.py:30-60 — transcribed in
tests/test_oracle_golden.py).  Floating-point outputs are **parity unpinned** at
the mlx boundary (mlx is not installable here); they are cross-checked in fp32
against HuggingFace `transformers` Qwen2VLForConditionalGeneration with the same
weights (tests/test_oracle_vs_hf.py), which pins the wiring but not the bf16
rounding points.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import mlx_semantics as S
from .mlx_semantics import Rounder


# ---------------------------------------------------------------------------
# configs (reference models/qwen2_vl/config.py + HF config.json of the 2B ckpt)
# ---------------------------------------------------------------------------
@dataclass
class VisionCfg:
    depth: int = 32
    embed_dim: int = 1280
    hidden_size: int = 1536  # merger output dim == text hidden
    num_heads: int = 16
    patch_size: int = 14
    mlp_ratio: float = 4.0
    in_channels: int = 3
    spatial_merge_size: int = 2
    temporal_patch_size: int = 2
    layer_norm_eps: float = 1e-6


@dataclass
class TextCfg:
    hidden_size: int = 1536
    num_hidden_layers: int = 28
    intermediate_size: int = 8960
    num_attention_heads: int = 12
    num_key_value_heads: int = 2
    rms_norm_eps: float = 1e-6
    vocab_size: int = 151936
    rope_theta: float = 1000000.0
    mrope_section: Tuple[int, int, int] = (16, 24, 24)
    tie_word_embeddings: bool = True


@dataclass
class Cfg:
    text: TextCfg = field(default_factory=TextCfg)
    vision: VisionCfg = field(default_factory=VisionCfg)
    image_token_id: int = 151655
    video_token_id: int = 151656
    vision_start_token_id: int = 151652
    vision_end_token_id: int = 151653


def qwen2_vl_2b() -> Cfg:
    return Cfg()


def tiny_cfg(**kw) -> Cfg:
    """Small config with the same structure (GQA, mrope sections, merge 2x2)."""
    t = TextCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512,
                num_attention_heads=4, num_key_value_heads=2, vocab_size=1024,
                mrope_section=(8, 12, 12), tie_word_embeddings=True)
    v = VisionCfg(depth=2, embed_dim=160, hidden_size=256, num_heads=2)
    c = Cfg(text=t, vision=v, image_token_id=1000, video_token_id=1001,
            vision_start_token_id=1002, vision_end_token_id=1003)
    for k, val in kw.items():
        setattr(c, k, val)
    return c


# ---------------------------------------------------------------------------
# weights — names are the reference's post-`sanitize` names (qwen2_vl.py:179-190)
# ---------------------------------------------------------------------------
def weight_shapes(cfg: Cfg) -> Dict[str, Tuple[int, ...]]:
    t, v = cfg.text, cfg.vision
    H, I = t.hidden_size, t.intermediate_size
    hd = H // t.num_attention_heads
    kvd = t.num_key_value_heads * hd
    E = v.embed_dim
    Em = int(E * v.mlp_ratio)
    shapes: Dict[str, Tuple[int, ...]] = {}
    # HF layout [out, C, T, ps, ps]; the reference stores [out,T,ps,ps,C] after
    # vision.sanitize — both describe the same (out, C*T*ps*ps) GEMM when the
    # pixel row is ordered (C,T,ps,ps)  (vision.py:92-98).
    shapes["vision_tower.patch_embed.proj.weight"] = (
        E, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)
    for i in range(v.depth):
        p = f"vision_tower.blocks.{i}."
        shapes[p + "norm1.weight"] = (E,)
        shapes[p + "norm1.bias"] = (E,)
        shapes[p + "norm2.weight"] = (E,)
        shapes[p + "norm2.bias"] = (E,)
        shapes[p + "attn.qkv.weight"] = (3 * E, E)
        shapes[p + "attn.qkv.bias"] = (3 * E,)
        shapes[p + "attn.proj.weight"] = (E, E)
        shapes[p + "attn.proj.bias"] = (E,)
        shapes[p + "mlp.fc1.weight"] = (Em, E)
        shapes[p + "mlp.fc1.bias"] = (Em,)
        shapes[p + "mlp.fc2.weight"] = (E, Em)
        shapes[p + "mlp.fc2.bias"] = (E,)
    m = v.spatial_merge_size ** 2 * E
    shapes["vision_tower.merger.ln_q.weight"] = (E,)
    shapes["vision_tower.merger.ln_q.bias"] = (E,)
    shapes["vision_tower.merger.mlp.0.weight"] = (m, m)
    shapes["vision_tower.merger.mlp.0.bias"] = (m,)
    shapes["vision_tower.merger.mlp.2.weight"] = (v.hidden_size, m)
    shapes["vision_tower.merger.mlp.2.bias"] = (v.hidden_size,)
    shapes["language_model.model.embed_tokens.weight"] = (t.vocab_size, H)
    for i in range(t.num_hidden_layers):
        p = f"language_model.model.layers.{i}."
        shapes[p + "input_layernorm.weight"] = (H,)
        shapes[p + "self_attn.q_proj.weight"] = (H, H)
        shapes[p + "self_attn.q_proj.bias"] = (H,)
        shapes[p + "self_attn.k_proj.weight"] = (kvd, H)
        shapes[p + "self_attn.k_proj.bias"] = (kvd,)
        shapes[p + "self_attn.v_proj.weight"] = (kvd, H)
        shapes[p + "self_attn.v_proj.bias"] = (kvd,)
        shapes[p + "self_attn.o_proj.weight"] = (H, H)
        shapes[p + "post_attention_layernorm.weight"] = (H,)
        shapes[p + "mlp.gate_proj.weight"] = (I, H)
        shapes[p + "mlp.up_proj.weight"] = (I, H)
        shapes[p + "mlp.down_proj.weight"] = (H, I)
    shapes["language_model.model.norm.weight"] = (H,)
    if not t.tie_word_embeddings:
        shapes["language_model.lm_head.weight"] = (t.vocab_size, H)
    return shapes


def init_weights(cfg: Cfg, seed: int = 0, dtype: str = "bf16", std: float = 0.02,
                 norm_jitter: float = 0.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights (N(0, 0.02), norm weights 1).

    One generator stream per tensor (seed, index) so that the SAME values can be
    regenerated tensor by tensor anywhere (CPU here, and uploaded to the GPU by
    the tests) without holding everything twice.  `norm_jitter` perturbs norm
    weights/biases so tests exercise the multiply/add rounding points.
    Returned tensors are fp32 with values representable in `dtype`.
    """
    R = Rounder(dtype)
    out = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        is_norm = ("norm" in name or "ln_q" in name) and len(shape) == 1
        if is_norm and name.endswith("weight"):
            w = torch.ones(shape)
            if norm_jitter:
                w = w + norm_jitter * torch.randn(shape, generator=g)
        elif is_norm and name.endswith("bias"):
            w = torch.zeros(shape)
            if norm_jitter:
                w = w + norm_jitter * torch.randn(shape, generator=g)
        else:
            w = std * torch.randn(shape, generator=g)
        out[name] = R.r(w)
    return out


# ---------------------------------------------------------------------------
# host preprocessing (processing_qwen3_vl.py:182-205, 302-354)
# ---------------------------------------------------------------------------
OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def smart_resize(height, width, factor=28, min_pixels=56 * 56,
                 max_pixels=14 * 14 * 4 * 1280):
    if max(height, width) / min(height, width) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def preprocess_image(img_chw_u8: np.ndarray, vcfg: VisionCfg,
                     mean=OPENAI_CLIP_MEAN, std=OPENAI_CLIP_STD):
    """(C,H,W) uint8 -> (grid_h*grid_w, C*tps*ps*ps) f32 rows in merge-group-major
    order, plus [1, grid_h, grid_w].  Sizes that smart_resize leaves unchanged
    take no resample (processing_qwen3_vl.py:164-170)."""
    C, H, W = img_chw_u8.shape
    ps, tps, ms = vcfg.patch_size, vcfg.temporal_patch_size, vcfg.spatial_merge_size
    rh, rw = smart_resize(H, W, factor=ps * ms)
    if (rh, rw) != (H, W):
        from PIL import Image
        pil = Image.fromarray(np.transpose(img_chw_u8, (1, 2, 0)))
        pil = pil.resize((rw, rh), resample=Image.BICUBIC)
        img_chw_u8 = np.transpose(np.array(pil), (2, 0, 1))
    img = img_chw_u8.astype(np.float32) * np.float32(1 / 255.0)
    m = np.array(mean, dtype=np.float32)[:, None, None]
    s = np.array(std, dtype=np.float32)[:, None, None]
    img = (img - m) / s
    patches = np.repeat(img[None, None, ...], tps, axis=1)
    gh, gw = rh // ps, rw // ps
    patches = patches.reshape(1, 1, tps, C, gh // ms, ms, ps, gw // ms, ms, ps)
    patches = patches.transpose(0, 1, 4, 7, 5, 8, 3, 2, 6, 9)
    flat = patches.reshape(gh * gw, C * tps * ps * ps)
    return np.ascontiguousarray(flat), [1, gh, gw]


def expand_image_tokens(ids: Sequence[int], grids: Sequence[Sequence[int]], cfg: Cfg):
    """processing_qwen2_vl.py:93-105: each image_token -> grid.prod()//merge^2 copies."""
    out, gi = [], 0
    m2 = cfg.vision.spatial_merge_size ** 2
    for t in ids:
        if t == cfg.image_token_id:
            n = int(np.prod(grids[gi])) // m2
            out.extend([cfg.image_token_id] * n)
            gi += 1
        else:
            out.append(int(t))
    return out


# ---------------------------------------------------------------------------
# get_rope_index  (language.py:216-402)  — integer, bit-exact
# ---------------------------------------------------------------------------
def get_rope_index(cfg: Cfg, input_ids, image_grid_thw=None, video_grid_thw=None,
                   attention_mask=None):
    """input_ids (B,T) ints; grids (n,3) ints -> (position_ids, deltas).

    With grids: position_ids (3,B,T), deltas (B,1).  Without: the text-only
    branch (:379-402) returning 2-D (B,T) position ids.
    """
    ids = np.asarray(input_ids, dtype=np.int64)
    B, T = ids.shape
    ms = cfg.vision.spatial_merge_size
    img_id, vid_id, vs_id = cfg.image_token_id, cfg.video_token_id, cfg.vision_start_token_id
    if image_grid_thw is not None or video_grid_thw is not None:
        mask = (np.ones_like(ids) if attention_mask is None
                else np.asarray(attention_mask, dtype=np.int64))
        pos = np.ones((3, B, T), dtype=np.int64)
        deltas: List[int] = []
        ii = vi = 0
        for i in range(B):
            row_mask = mask[i].tolist()
            toks = [t for t, k in zip(ids[i].tolist(), row_mask) if k == 1]
            vision_tokens = [toks[j + 1] for j, t in enumerate(toks[:-1]) if t == vs_id]
            n_img = sum(t == img_id for t in vision_tokens)
            n_vid = sum(t == vid_id for t in vision_tokens)
            chunks: List[np.ndarray] = []
            st = 0
            rem_i, rem_v = n_img, n_vid
            for _ in range(n_img + n_vid):
                ed_i = toks.index(img_id, st) if (img_id in toks and rem_i > 0) else len(toks) + 1
                ed_v = toks.index(vid_id, st) if (vid_id in toks and rem_v > 0) else len(toks) + 1
                if ed_i < ed_v:
                    t, h, w = (int(x) for x in image_grid_thw[ii])
                    ii += 1
                    rem_i -= 1
                    ed = ed_i
                else:
                    t, h, w = (int(x) for x in video_grid_thw[vi])
                    vi += 1
                    rem_v -= 1
                    ed = ed_v
                gt, gh, gw = t, h // ms, w // ms
                text_len = ed - st
                st_idx = int(chunks[-1].max()) + 1 if chunks else 0
                chunks.append(np.broadcast_to(np.arange(text_len)[None], (3, text_len)) + st_idx)
                ti = np.broadcast_to(np.arange(gt)[:, None], (gt, gh * gw)).reshape(-1)
                hi = np.broadcast_to(np.arange(gh)[None, :, None], (gt, gh, gw)).reshape(-1)
                wi = np.broadcast_to(np.arange(gw)[None, None, :], (gt, gh, gw)).reshape(-1)
                chunks.append(np.stack([ti, hi, wi]) + text_len + st_idx)
                st = ed + gt * gh * gw
            if st < len(toks):
                st_idx = int(chunks[-1].max()) + 1 if chunks else 0
                text_len = len(toks) - st
                chunks.append(np.broadcast_to(np.arange(text_len)[None], (3, text_len)) + st_idx)
            if not chunks:
                deltas.append(0)
                continue
            llm = np.concatenate(chunks, axis=1).reshape(3, -1)
            cmax = int(llm.max())
            ci = 0
            for col, keep in enumerate(row_mask):
                if keep == 1:
                    pos[:, i, col] = llm[:, ci]
                    ci += 1
            deltas.append(cmax + 1 - len(toks))
        return pos, np.asarray(deltas, dtype=np.int64).reshape(-1, 1)
    if attention_mask is not None:
        mask = np.asarray(attention_mask, dtype=np.int64)
        pos = np.cumsum(mask, axis=-1) - 1
        pos = np.where(mask == 0, 1, pos)
        mx_ = pos.max(axis=-1, keepdims=True)
        return pos, mx_ + 1 - mask.shape[-1]
    pos = np.broadcast_to(np.arange(T)[None], (B, T)).copy()
    return pos, np.zeros((B, 1), dtype=np.int64)


# ---------------------------------------------------------------------------
# merge_input_ids_with_image_features  (qwen2_vl.py:78-148) — indexing bit-exact
# ---------------------------------------------------------------------------
def merge_indices(cfg: Cfg, input_ids) -> np.ndarray:
    """Returns src (B,T) int64: src[b,t] = row of image_features that replaces
    position (b,t), or -1 where the text embedding is kept.  Raises ValueError
    exactly where the reference does (count mismatch is checked by the caller
    that knows n_features, see merge_input_ids_with_image_features)."""
    ids = np.asarray(input_ids, dtype=np.int64)
    mask = ids == cfg.image_token_id
    if mask.sum() == 0:
        mask = ids == cfg.video_token_id
    src = np.full(ids.shape, -1, dtype=np.int64)
    start = 0
    for b in range(ids.shape[0]):
        n = int(mask[b].sum())
        if n > 0:
            cs = np.cumsum(mask[b].astype(np.int32)) - 1
            src[b] = np.where(mask[b], cs + start, -1)
            start += n
    return src


def merge_input_ids_with_image_features(cfg: Cfg, image_features, inputs_embeds, input_ids):
    ids = np.asarray(input_ids, dtype=np.int64)
    mask = ids == cfg.image_token_id
    if mask.sum() == 0:
        mask = ids == cfg.video_token_id
    start = 0
    outs = []
    for b in range(ids.shape[0]):
        n = int(mask[b].sum())
        if n > 0:
            feats = image_features[start:start + n]
            if feats.shape[0] != n:
                raise ValueError(
                    f"Number of image token positions ({n}) does not match "
                    f"number of image features ({feats.shape[0]}) for batch {b}")
            cs = np.cumsum(mask[b].astype(np.int32))
            idx = torch.from_numpy(np.where(mask[b], cs - 1, 0))
            gathered = feats[idx]
            m = torch.from_numpy(mask[b])[:, None]
            outs.append(torch.where(m, gathered, inputs_embeds[b]))
            start += n
        else:
            outs.append(inputs_embeds[b])
    return torch.stack(outs, 0)


# ---------------------------------------------------------------------------
# Vision tower (vision.py)
# ---------------------------------------------------------------------------
def rot_pos_ids(grid_thw, ms: int) -> np.ndarray:
    """vision.py:219-249: (N,2) [h,w] ids per patch in merge-group-major order."""
    out = []
    for t, h, w in grid_thw:
        t, h, w = int(t), int(h), int(w)
        hp = np.repeat(np.arange(h)[:, None], w, axis=1)
        hp = hp.reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        wp = np.repeat(np.arange(w)[None, :], h, axis=0)
        wp = wp.reshape(h // ms, ms, w // ms, ms).transpose(0, 2, 1, 3).reshape(-1)
        out.append(np.tile(np.stack([hp, wp], axis=-1), (t, 1)))
    return np.concatenate(out, axis=0)


def vision_rotary_freqs(grid_thw, vcfg: VisionCfg) -> torch.Tensor:
    """vision.py:53-65, 219-255: (N, head_dim/2) fp32 angles [h-freqs | w-freqs]."""
    hd = vcfg.embed_dim // vcfg.num_heads
    dim = hd // 2
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
    g = np.asarray(grid_thw)
    max_grid = int(g[:, 1:].max())
    seq = torch.arange(max_grid, dtype=torch.float32)
    full = torch.outer(seq, inv_freq)  # (max_grid, dim/2)
    pid = torch.from_numpy(rot_pos_ids(grid_thw, vcfg.spatial_merge_size))
    return full[pid].reshape(pid.shape[0], -1)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def vision_forward(cfg: Cfg, W: Dict[str, torch.Tensor], pixel_values, grid_thw,
                   R: Rounder, return_blocks: bool = False):
    """pixel_values (N, C*tps*ps*ps) f32 -> merged features (N/4, hidden)."""
    v = cfg.vision
    E, nh = v.embed_dim, v.num_heads
    hd = E // nh
    x = R.r(torch.as_tensor(pixel_values, dtype=torch.float32))  # qwen2_vl.py:44-45 astype
    # PatchEmbed: Conv3d kernel==stride, no bias == GEMM over (C,T,ps,ps)-ordered rows
    wpe = W["vision_tower.patch_embed.proj.weight"].reshape(E, -1)
    h = S.linear(R, x, wpe)
    freqs = vision_rotary_freqs(grid_thw, v)  # (N, hd/2) fp32
    cos = torch.cos(freqs).repeat(1, 2)[:, None, :]  # (N,1,hd) fp32
    sin = torch.sin(freqs).repeat(1, 2)[:, None, :]
    g = np.asarray(grid_thw)
    seg = []
    for i in range(g.shape[0]):
        seg += [int(g[i, 1] * g[i, 2])] * int(g[i, 0])
    cu = np.concatenate([[0], np.cumsum(seg)])
    scale = hd ** -0.5
    blocks_out = []
    for i in range(v.depth):
        p = f"vision_tower.blocks.{i}."
        y = S.layer_norm(R, h, W[p + "norm1.weight"], W[p + "norm1.bias"], v.layer_norm_eps)
        qkv = S.linear(R, y, W[p + "attn.qkv.weight"], W[p + "attn.qkv.bias"])
        N = qkv.shape[0]
        qkv = qkv.reshape(N, 3, nh, hd)
        q, k, vv = qkv[:, 0], qkv[:, 1], qkv[:, 2]  # (N, nh, hd)
        # apply_rotary_pos_emb_vision: act-dtype tensor * fp32 cos -> fp32; one cast
        q = R.r(q * cos + _rotate_half(q) * sin)
        k = R.r(k * cos + _rotate_half(k) * sin)
        outs = []
        for s in range(len(cu) - 1):
            a, b = int(cu[s]), int(cu[s + 1])
            qs = q[a:b].transpose(0, 1)[None]
            ks = k[a:b].transpose(0, 1)[None]
            vs = vv[a:b].transpose(0, 1)[None]
            o = S.sdpa(R, qs, ks, vs, scale, causal=False)  # (1,nh,n,hd)
            outs.append(o[0].transpose(0, 1).reshape(b - a, E))
        att = torch.cat(outs, 0)
        att = S.linear(R, att, W[p + "attn.proj.weight"], W[p + "attn.proj.bias"])
        h = R.r(h + att)
        y = S.layer_norm(R, h, W[p + "norm2.weight"], W[p + "norm2.bias"], v.layer_norm_eps)
        y = S.linear(R, y, W[p + "mlp.fc1.weight"], W[p + "mlp.fc1.bias"])
        y = S.gelu_fast(R, y)
        y = S.linear(R, y, W[p + "mlp.fc2.weight"], W[p + "mlp.fc2.bias"])
        h = R.r(h + y)
        if return_blocks:
            blocks_out.append(h)
    y = S.layer_norm(R, h, W["vision_tower.merger.ln_q.weight"],
                     W["vision_tower.merger.ln_q.bias"], 1e-6)
    y = y.reshape(-1, E * v.spatial_merge_size ** 2)
    y = S.linear(R, y, W["vision_tower.merger.mlp.0.weight"], W["vision_tower.merger.mlp.0.bias"])
    y = S.gelu_exact(R, y)
    y = S.linear(R, y, W["vision_tower.merger.mlp.2.weight"], W["vision_tower.merger.mlp.2.bias"])
    if return_blocks:
        return y, blocks_out
    return y


# ---------------------------------------------------------------------------
# M-RoPE (rope_utils.py)
# ---------------------------------------------------------------------------
def mrope_selector(section: Sequence[int], freq_dim: int) -> np.ndarray:
    """_chunked_position_selector (rope_utils.py:519-526)."""
    sel = [0] * freq_dim
    off = section[0]
    for dim, length in enumerate(section[1:], start=1):
        for idx in range(off, min(off + length, freq_dim)):
            sel[idx] = dim
        off += length
    return np.asarray(sel, dtype=np.int64)


def mrope_cos_sin(tcfg: TextCfg, position_ids, R: Rounder):
    """MRoPERotaryEmbedding.__call__ (rope_utils.py:1227-1241), cast_output=True:
    fp32 angles, cos/sin cast to the activation dtype.  position_ids (3,B,L) or
    (B,L) ints -> cos, sin (B, L, head_dim)."""
    hd = tcfg.hidden_size // tcfg.num_attention_heads
    inv_freq = 1.0 / (tcfg.rope_theta ** (torch.arange(0, hd, 2).to(torch.float32) / hd))
    pos = torch.as_tensor(np.asarray(position_ids))
    if pos.ndim == 2:
        freqs = pos.to(torch.float32)[..., None] * inv_freq
    else:
        sel = torch.from_numpy(mrope_selector(tcfg.mrope_section, inv_freq.shape[0]))
        p = pos[sel].permute(1, 2, 0)  # (B,L,F): slot j takes axis sel[j]
        freqs = p.to(torch.float32) * inv_freq
    emb = torch.cat([freqs, freqs], dim=-1)
    return R.r(torch.cos(emb)), R.r(torch.sin(emb))


def apply_mrope(R: Rounder, q, k, cos, sin):
    """apply_multimodal_rotary_pos_emb(style="chunked") -> _apply_rotary_embedding
    (rope_utils.py:1301-1334) with no compute_dtype: three act-dtype roundings."""
    c, s = cos[:, None], sin[:, None]
    qe = R.r(R.r(q * c) + R.r(_rotate_half(q) * s))
    ke = R.r(R.r(k * c) + R.r(_rotate_half(k) * s))
    return qe, ke


# ---------------------------------------------------------------------------
# KVCache (cache.py:337-439)
# ---------------------------------------------------------------------------
class OracleKVCache:
    step = 256

    def __init__(self):
        self.keys = None
        self.values = None
        self.offset = 0

    def update_and_fetch(self, k, v):
        prev = self.offset
        if self.keys is None or (prev + k.shape[2]) > self.keys.shape[2]:
            B, nkv, _, hd = k.shape
            n_steps = (self.step + k.shape[2] - 1) // self.step
            nk = torch.zeros(B, nkv, n_steps * self.step, hd)
            nv = torch.zeros(B, nkv, n_steps * self.step, v.shape[3])
            if self.keys is not None:
                if prev % self.step != 0:
                    self.keys = self.keys[..., :prev, :]
                    self.values = self.values[..., :prev, :]
                self.keys = torch.cat([self.keys, nk], dim=2)
                self.values = torch.cat([self.values, nv], dim=2)
            else:
                self.keys, self.values = nk, nv
        self.offset += k.shape[2]
        self.keys[..., prev:self.offset, :] = k
        self.values[..., prev:self.offset, :] = v
        return self.keys[..., :self.offset, :], self.values[..., :self.offset, :]

    def trim(self, n):
        n = min(self.offset, n)
        self.offset -= n
        return n


# ---------------------------------------------------------------------------
# Language model (language.py)
# ---------------------------------------------------------------------------
def lm_layers_forward(cfg: Cfg, W, h, position_ids, cache: List[OracleKVCache], R: Rounder,
                      collect: Optional[list] = None, mask=None):
    """Qwen2Model.__call__ without the embedding: h (B,L,H) -> final-normed (B,L,H)."""
    t = cfg.text
    B, L, H = h.shape
    nh, nkv = t.num_attention_heads, t.num_key_value_heads
    hd = H // nh
    cos, sin = mrope_cos_sin(t, position_ids, R)
    scale = hd ** -0.5
    for i in range(t.num_hidden_layers):
        p = f"language_model.model.layers.{i}."
        x = S.rms_norm(R, h, W[p + "input_layernorm.weight"], t.rms_norm_eps)
        q = S.linear(R, x, W[p + "self_attn.q_proj.weight"], W[p + "self_attn.q_proj.bias"])
        k = S.linear(R, x, W[p + "self_attn.k_proj.weight"], W[p + "self_attn.k_proj.bias"])
        v = S.linear(R, x, W[p + "self_attn.v_proj.weight"], W[p + "self_attn.v_proj.bias"])
        q = q.reshape(B, L, nh, hd).transpose(1, 2)
        k = k.reshape(B, L, nkv, hd).transpose(1, 2)
        v = v.reshape(B, L, nkv, hd).transpose(1, 2)
        q, k = apply_mrope(R, q, k, cos, sin)
        keys, values = cache[i].update_and_fetch(k, v)
        o = S.sdpa(R, q, keys, values, scale, causal=(L > 1), mask=mask)
        o = o.transpose(1, 2).reshape(B, L, H)
        r = S.linear(R, o, W[p + "self_attn.o_proj.weight"])
        h = R.r(h + r)
        x = S.rms_norm(R, h, W[p + "post_attention_layernorm.weight"], t.rms_norm_eps)
        g = S.linear(R, x, W[p + "mlp.gate_proj.weight"])
        u = S.linear(R, x, W[p + "mlp.up_proj.weight"])
        d = S.linear(R, S.swiglu(R, g, u), W[p + "mlp.down_proj.weight"])
        h = R.r(h + d)
        if collect is not None:
            collect.append(h)
    return S.rms_norm(R, h, W["language_model.model.norm.weight"], t.rms_norm_eps)


def lm_head(cfg: Cfg, W, hidden, R: Rounder):
    w = (W["language_model.model.embed_tokens.weight"] if cfg.text.tie_word_embeddings
         else W["language_model.lm_head.weight"])
    return S.linear(R, hidden, w)


def decode_position_ids(cache_offset: int, rope_deltas, B: int):
    """language.py:476-509 for L==1: position = cache_offset + rope_delta, the same
    on all three M-RoPE axes."""
    d = np.asarray(rope_deltas, dtype=np.int64).reshape(-1)
    if d.shape[0] < B:
        d = np.tile(d, B)[:B]
    pos = (cache_offset + d[:B])[:, None]  # (B,1)
    return np.broadcast_to(pos[None], (3, B, 1)).copy()


@dataclass
class PrefillOut:
    input_ids: np.ndarray
    inputs_embeds: torch.Tensor
    image_features: Optional[torch.Tensor]
    position_ids: np.ndarray
    rope_deltas: np.ndarray
    logits_last: torch.Tensor  # (B, V) logits of the last prompt row


def get_input_embeddings(cfg: Cfg, W, input_ids, pixel_values, grid_thw, R: Rounder,
                         attention_mask=None):
    """Model.get_input_embeddings (qwen2_vl.py:20-76)."""
    ids = torch.as_tensor(np.asarray(input_ids, dtype=np.int64))
    emb = W["language_model.model.embed_tokens.weight"][ids]
    if pixel_values is None:
        pos, deltas = get_rope_index(cfg, input_ids, attention_mask=attention_mask)
        return emb, None, pos, deltas
    feats = vision_forward(cfg, W, pixel_values, grid_thw, R)
    merged = merge_input_ids_with_image_features(cfg, feats, emb, input_ids)
    pos, deltas = get_rope_index(cfg, input_ids, grid_thw, None, attention_mask)
    return merged, feats, pos, deltas


def logprobs_from_logits(R: Rounder, logits):
    """ar.py:368: logits - logsumexp(logits), in the logits dtype."""
    return R.r(logits - S.logsumexp(R, logits))


def greedy_generate(cfg: Cfg, W, input_ids, pixel_values, grid_thw, max_tokens: int,
                    dtype: str = "bf16", force_tokens: Optional[Sequence[int]] = None,
                    keep_logits: bool = True):
    """generate_step (ar.py:151-515) for the greedy sampler, EOS ignored.

    Returns dict(tokens, logits [per step (V,)], prefill: PrefillOut, cache).
    `force_tokens` teacher-forces the fed-back token (for per-step logits parity
    that does not depend on argmax ties).
    """
    R = Rounder(dtype)
    t = cfg.text
    ids = np.asarray(input_ids, dtype=np.int64)
    B, T = ids.shape
    embeds, feats, pos, deltas = get_input_embeddings(cfg, W, ids, pixel_values, grid_thw, R)
    cache = [OracleKVCache() for _ in range(t.num_hidden_layers)]
    hidden = lm_layers_forward(cfg, W, embeds, pos, cache, R)
    logits = lm_head(cfg, W, hidden[:, -1, :], R)  # reference computes all rows, slices [-1]
    pre = PrefillOut(ids, embeds, feats, np.asarray(pos), np.asarray(deltas), logits)
    toks, all_logits, all_lp = [], [], []
    for n in range(max_tokens):
        lp = logprobs_from_logits(R, logits)
        y = S.argmax_lowest(lp)  # (B,)
        toks.append(y.clone())
        if keep_logits:
            all_logits.append(logits.clone())
            all_lp.append(lp)
        if n == max_tokens - 1:
            break
        feed = y if force_tokens is None else torch.full_like(y, int(force_tokens[n]))
        e = W["language_model.model.embed_tokens.weight"][feed][:, None, :]
        p = decode_position_ids(cache[0].offset, deltas, B)
        hidden = lm_layers_forward(cfg, W, e, p, cache, R)
        logits = lm_head(cfg, W, hidden[:, -1, :], R)
    return dict(tokens=torch.stack(toks, 1), logits=all_logits, logprobs=all_lp,
                prefill=pre, cache=cache)


# ---------------------------------------------------------------------------
# synthetic request
# ---------------------------------------------------------------------------
def synthetic_request(cfg: Cfg, n_text: int, image_hw=(336, 336), seed: int = 0,
                      text_vocab: Optional[int] = None):
    """uint8 random image + prompt of `n_text` text tokens (incl. vision_start /
    vision_end) with one image placeholder expanded by the processor rule."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(3, image_hw[0], image_hw[1]), dtype=np.uint8)
    pv, grid = preprocess_image(img, cfg.vision)
    hi = text_vocab or min(cfg.text.vocab_size, cfg.image_token_id) - 16
    text = rng.integers(0, max(hi, 8), size=n_text).tolist()
    n_pre = min(4, n_text // 2)
    ids = text[:n_pre] + [cfg.vision_start_token_id, cfg.image_token_id,
                          cfg.vision_end_token_id] + text[n_pre + 2:]
    ids = ids[:n_text + 1]  # n_text text-side tokens + 1 placeholder
    ids = expand_image_tokens(ids, [grid], cfg)
    return dict(image=img, pixel_values=pv, image_grid_thw=np.asarray([grid], dtype=np.int64),
                input_ids=np.asarray([ids], dtype=np.int64))
