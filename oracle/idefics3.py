"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's Idefics3 / SmolVLM generate path (SmolVLM is an alias,
models/smolvlm/smolvlm.py):
  models/idefics3/vision.py:67-150   VisionEmbeddings: Conv2d 14x14 + bias on fp32 pixels, position ids = buckets of the
                                     fractional coordinates of the VALID block (sum(frac >= boundaries): the correct
                                     bucketing, unlike Idefics2's digitize - 1) written to the FIRST n_valid sequence
                                     positions (not to the valid positions — reproduced), position embeddings zeroed on
                                     padding patches
  models/idefics3/vision.py:152-185  VisionModel: the embeddings are CAST TO THE WEIGHT DTYPE (bf16) before the encoder —
                                     unlike LLaVA / Idefics2 this tower runs in bf16 —, N x {LN, MHA with bias, LN,
                                     GELU(approx="precise") MLP}, post-LayerNorm with eps 1e-5
  models/idefics3/idefics3.py:47-70  connector: pixel shuffle (scale_factor) + Linear without bias, in fp32
                                     (`pooler_output.astype(pixel_values.dtype)`)
  models/idefics3/idefics3.py:80-175 padding-image removal, pixel mask -> patch mask, masked-scatter merge (one rounding)
  models/idefics3/language.py        Llama decoder (== the Qwen2 layer without q/k/v bias, 1-D rotary positions)
The integer logic is pinned by executing the reference's own source (tests/golden/make_idefics3_golden.py); the
floating-point rounding points are those of oracle/mlx_semantics.py (unpinned at the mlx boundary, stated there)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import idefics2 as I2
from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder


@dataclass
class Idefics3Cfg:
    vision: I2.SiglipCfg = field(default_factory=lambda: I2.SiglipCfg(hidden_size=1152, num_hidden_layers=27,
                                                                       intermediate_size=4304, num_attention_heads=16,
                                                                       image_size=364, patch_size=14))
    text: I2.MistralCfg = field(default_factory=lambda: I2.MistralCfg(vocab_size=128259, rope_theta=500000.0))
    scale_factor: int = 2
    image_token_index: int = 128257


def pixel_shuffle(x: torch.Tensor, s: int) -> torch.Tensor:
    """idefics3.py:47-62.  (B, seq, E) -> (B, seq / s^2, E s^2): output row (yg, xg), chunk (dy, dx) = source token
    (yg s + dy, xg s + dx)"""
    B, seq, E = x.shape
    side = int(seq ** 0.5)
    x = x.reshape(B, side, side, E).reshape(B, side, side // s, E * s).transpose(1, 2)
    x = x.reshape(B, side // s, side // s, E * s * s).transpose(1, 2)
    return x.reshape(B, seq // (s * s), E * s * s)


def mlx_arange_f32(start: float, stop: float, step: float) -> np.ndarray:
    """mx.arange(start, stop, step) with python floats: a float32 array of ceil((stop - start) / step) elements, element i
    = start + i * step evaluated in float32 with the product rounded (mlx's Metal kernel `out[i] = start + i * step`, the
    device the reference runs on; scalars are the float32 casts of start and of (start + step) - start).  PARITY NOTE: the
    position buckets below compare k / n against these boundaries with `>=`, and for an image that fills the grid every
    coordinate is a TIE, so the ids depend on this rounding (float64 boundaries, an FMA, or the CPU backend's running sum
    each move a few ids by one); which one a given mlx build produces is unpinned — HF's torch.bucketize gives the
    identity there."""
    n = max(int(np.ceil((stop - start) / step)), 0)
    s0 = np.float32(start)
    st = np.float32(np.float32(start + step) - s0)
    return np.array([np.float32(s0 + np.float32(np.float32(i) * st)) for i in range(n)], dtype=np.float32)


def position_ids_and_mask(patch_mask: Optional[np.ndarray], gh: int, gw: int, side: int):
    """vision.py:95-141 -> (ids (B, gh*gw) int64, mask (B, gh*gw) bool or None)"""
    if patch_mask is None:
        return None, None
    m = np.asarray(patch_mask).astype(bool)
    B = m.shape[0]
    seq = gh * gw
    bounds = mlx_arange_f32(1 / side, 1.0, 1 / side)
    hi = np.float32(1.0 - 1e-6)
    ids = np.zeros((B, seq), dtype=np.int64)
    for b in range(B):
        nh = max(int(m[b, :, 0].sum()), 1)
        nw = max(int(m[b, 0, :].sum()), 1)
        fh = np.clip(np.arange(nh, dtype=np.float32) / np.float32(nh), np.float32(0.0), hi)
        fw = np.clip(np.arange(nw, dtype=np.float32) / np.float32(nw), np.float32(0.0), hi)
        bh = (fh[:, None] >= bounds[None, :]).sum(axis=1)
        bw = (fw[:, None] >= bounds[None, :]).sum(axis=1)
        p = (bh[:, None] * side + bw[None, :]).reshape(-1)
        n = min(p.shape[0], seq)
        ids[b, :n] = p[:n]
    return ids, m.reshape(B, -1)[:, :seq]


def weight_shapes(cfg: Idefics3Cfg) -> Dict[str, Tuple[int, ...]]:
    v, t = cfg.vision, cfg.text
    E, I, H = v.hidden_size, v.intermediate_size, t.hidden_size
    s: Dict[str, Tuple[int, ...]] = {}
    p = "vision_model."
    s[p + "embeddings.patch_embedding.weight"] = (E, v.patch_size, v.patch_size, v.num_channels)
    s[p + "embeddings.patch_embedding.bias"] = (E,)
    s[p + "embeddings.position_embedding.weight"] = ((v.image_size // v.patch_size) ** 2, E)
    for i in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            s[q + n + ".weight"], s[q + n + ".bias"] = (E,), (E,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[q + f"self_attn.{n}.weight"], s[q + f"self_attn.{n}.bias"] = (E, E), (E,)
        s[q + "mlp.fc1.weight"], s[q + "mlp.fc1.bias"] = (I, E), (I,)
        s[q + "mlp.fc2.weight"], s[q + "mlp.fc2.bias"] = (E, I), (E,)
    s[p + "post_layernorm.weight"], s[p + "post_layernorm.bias"] = (E,), (E,)
    s["connector.modality_projection.proj.weight"] = (H, E * cfg.scale_factor ** 2)
    hd = H // t.num_attention_heads
    lkv = t.num_key_value_heads * hd
    s["language_model.embed_tokens.weight"] = (t.vocab_size, H)
    for i in range(t.num_hidden_layers):
        q = f"language_model.layers.{i}."
        s[q + "input_layernorm.weight"] = (H,)
        s[q + "post_attention_layernorm.weight"] = (H,)
        s[q + "self_attn.q_proj.weight"] = (H, H)
        s[q + "self_attn.k_proj.weight"] = (lkv, H)
        s[q + "self_attn.v_proj.weight"] = (lkv, H)
        s[q + "self_attn.o_proj.weight"] = (H, H)
        s[q + "mlp.gate_proj.weight"] = (t.intermediate_size, H)
        s[q + "mlp.up_proj.weight"] = (t.intermediate_size, H)
        s[q + "mlp.down_proj.weight"] = (H, t.intermediate_size)
    s["language_model.norm.weight"] = (H,)
    s["language_model.lm_head.weight"] = (t.vocab_size, H)
    return s


def init_weights(cfg: Idefics3Cfg, seed: int = 0, std: float = 0.02, norm_jitter: float = 0.05):
    W = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        if name.endswith(".weight") and ("norm" in name.split(".")[-2]):
            x = 1.0 + norm_jitter * torch.randn(shape, generator=g)
        else:
            x = std * torch.randn(shape, generator=g)
        W[name] = x.to(torch.bfloat16).to(torch.float32)
    return W


def vision_forward(cfg: Idefics3Cfg, W, pixel_values_nhwc: torch.Tensor, patch_mask: Optional[np.ndarray], dtype: str = "bf16"):
    """-> pooler output (n_img, gh*gw, E) in the encoder's dtype (bf16 values)"""
    v = cfg.vision
    p = "vision_model."
    F, R = Rounder("f32"), Rounder(dtype)
    x = pixel_values_nhwc.to(torch.float32)
    B, Hh, Ww, C = x.shape
    ps = v.patch_size
    gh, gw = Hh // ps, Ww // ps
    patches = x.reshape(B, gh, ps, gw, ps, C).permute(0, 1, 3, 2, 4, 5).reshape(B, gh * gw, ps * ps * C)
    emb = S.linear(F, patches, W[p + "embeddings.patch_embedding.weight"].reshape(v.hidden_size, -1),
                   W[p + "embeddings.patch_embedding.bias"])                       # fp32: float32 pixels x bf16 weights
    table = W[p + "embeddings.position_embedding.weight"]
    ids, m = position_ids_and_mask(patch_mask, gh, gw, v.image_size // ps)
    if ids is None:
        pos = table[torch.arange(gh * gw)][None].expand(B, -1, -1)
    else:
        pos = R.r(table[torch.from_numpy(ids)] * torch.from_numpy(m)[..., None].to(torch.float32))  # bf16 x bool -> bf16
    h = R.r(emb + pos)                                                              # fp32 sum, then astype(weight dtype)
    nh = v.num_attention_heads
    hd = v.hidden_size // nh
    for i in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{i}."
        y = S.layer_norm(R, h, W[q + "layer_norm1.weight"], W[q + "layer_norm1.bias"], v.layer_norm_eps)
        L = y.shape[1]
        qq = S.linear(R, y, W[q + "self_attn.q_proj.weight"], W[q + "self_attn.q_proj.bias"]).reshape(B, L, nh, hd).transpose(1, 2)
        kk = S.linear(R, y, W[q + "self_attn.k_proj.weight"], W[q + "self_attn.k_proj.bias"]).reshape(B, L, nh, hd).transpose(1, 2)
        vv = S.linear(R, y, W[q + "self_attn.v_proj.weight"], W[q + "self_attn.v_proj.bias"]).reshape(B, L, nh, hd).transpose(1, 2)
        o = S.sdpa(R, qq, kk, vv, hd ** -0.5, causal=False).transpose(1, 2).reshape(B, L, v.hidden_size)
        o = S.linear(R, o, W[q + "self_attn.out_proj.weight"], W[q + "self_attn.out_proj.bias"])
        h = R.r(h + o)
        y = S.layer_norm(R, h, W[q + "layer_norm2.weight"], W[q + "layer_norm2.bias"], v.layer_norm_eps)
        y = S.gelu_tanh(R, S.linear(R, y, W[q + "mlp.fc1.weight"], W[q + "mlp.fc1.bias"]))
        y = S.linear(R, y, W[q + "mlp.fc2.weight"], W[q + "mlp.fc2.bias"])
        h = R.r(h + y)
    return S.layer_norm(R, h, W[p + "post_layernorm.weight"], W[p + "post_layernorm.bias"], 1e-5)


def image_features(cfg: Idefics3Cfg, W, pixel_values_bnchw, pixel_attention_mask, dtype: str = "bf16") -> torch.Tensor:
    """-> (n_real * P / s^2, hidden) fp32 (the merge rounds them)"""
    pv = np.asarray(pixel_values_bnchw, dtype=np.float32)
    B, N, C, Hh, Ww = pv.shape
    keep = I2.real_image_indices(pv)
    pv = pv.reshape(B * N, C, Hh, Ww)[keep]
    if pixel_attention_mask is None:
        pam = np.ones((pv.shape[0], Hh, Ww), dtype=bool)
    else:
        pam = np.asarray(pixel_attention_mask).reshape(B * N, Hh, Ww)[keep]
    pmask = I2.patch_attention_mask(pam, cfg.vision.patch_size)
    pooled = vision_forward(cfg, W, torch.from_numpy(pv).permute(0, 2, 3, 1), pmask, dtype)
    x = pixel_shuffle(pooled.to(torch.float32), cfg.scale_factor)
    y = S.linear(Rounder("f32"), x, W["connector.modality_projection.proj.weight"])
    return y.reshape(-1, y.shape[-1])


def merge(cfg: Idefics3Cfg, image_features_: torch.Tensor, inputs_embeds: torch.Tensor, input_ids):
    return I2.merge(types_cfg(cfg), image_features_, inputs_embeds, input_ids)


def types_cfg(cfg):
    import types
    return types.SimpleNamespace(image_token_index=cfg.image_token_index)


def _as_qwen(cfg: Idefics3Cfg, W):
    c2 = I2.Idefics2Cfg(vision=cfg.vision, text=cfg.text, image_token_index=cfg.image_token_index)
    return I2._as_qwen(c2, W)


def greedy_generate(cfg: Idefics3Cfg, W, input_ids, pixel_values_bnchw, pixel_attention_mask, max_tokens: int,
                    dtype: str = "bf16"):
    R = Rounder(dtype)
    qc, W2 = _as_qwen(cfg, W)
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    embeds = W["language_model.embed_tokens.weight"][ids]
    feats = None
    if pixel_values_bnchw is not None:
        feats = R.r(image_features(cfg, W, pixel_values_bnchw, pixel_attention_mask, dtype))
        embeds = merge(cfg, feats, embeds, input_ids)
    T = embeds.shape[1]
    cache = [Q.OracleKVCache() for _ in range(cfg.text.num_hidden_layers)]
    hidden = Q.lm_layers_forward(qc, W2, embeds, I2._positions(0, T), cache, R)
    logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
    out_logits, toks = [logits], []
    for n in range(max_tokens):
        y = S.argmax_lowest(Q.logprobs_from_logits(R, logits))
        toks.append(int(y[0]))
        if n == max_tokens - 1:
            break
        e = W["language_model.embed_tokens.weight"][y][:, None, :]
        hidden = Q.lm_layers_forward(qc, W2, e, I2._positions(cache[0].offset, 1), cache, R)
        logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
        out_logits.append(logits)
    return {"tokens": toks, "logits": out_logits, "image_features": feats, "inputs_embeds": embeds}
