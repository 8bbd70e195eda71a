"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's LLaVA-1.5 generate path:
  models/llava/vision.py:108-185   CLIP ViT-L/14 tower (Conv2d patch embedding without bias,
                                   class token, learned positions, pre-LayerNorm, N x {LN, MHA
                                   with bias, LN, fast-GELU MLP}, all hidden states returned)
  models/llava/llava.py:14-29      projector Linear -> GELU(exact) -> Linear
  models/llava/llava.py:33-88      get_input_embeddings: hidden_states[-2], drop CLS ("default")
  models/llava/llava.py:90-116     merge: inputs_embeds[:, positions of <image>, :] = features
  models/llava/language.py:16-150  Llama decoder with nn.RoPE(traditional=False) and KVCache

PARITY STATUS: floating-point rounding points unpinned at the mlx boundary (mlx is not installable offline; see
oracle/mlx_semantics.py).  The product path (mlx_vlm_b200/models/llava/) is checked against this file on the GPU
(tests/test_llava_gpu.py).  The structure is pinned two ways (tests/test_oracle_llava.py): the merge against the
reference's own function source (tests/golden/), and the whole model in fp32 against
HuggingFace transformers' LlavaForConditionalGeneration with the same weights.

Precision of the vision tower — what the reference actually computes: `prepare_inputs` builds
`pixel_values` as a float32 array (utils.py:2091) and LLaVA never casts it (llava.py:61-63;
only Qwen2-VL casts to the weight dtype, qwen2_vl.py:44-45).  mlx promotes float32 x bfloat16
to float32, so EVERY op of the tower and of the projector runs in fp32 with bf16-valued
weights; the features are rounded to the embedding dtype only at the merge (llava.py:101-104).
`vision_dtype="f32"` below is therefore the reference's semantics; "bf16" is kept to quantify
what a bf16 tower would change.

The language model reuses oracle/qwen2vl.py::lm_layers_forward: a Llama layer is a Qwen2 layer
without q/k/v bias, and nn.RoPE(traditional=False) is M-RoPE with the same position on all
three axes (half-split pairing, inv_freq = base^(-2i/d)); the rounding points of the rotary
step are those of that file (unpinned at the mlx boundary, stated there).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder


@dataclass
class ClipCfg:
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    intermediate_size: int = 4096
    num_attention_heads: int = 16
    image_size: int = 336
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-5

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2


@dataclass
class LlamaCfg:
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    intermediate_size: int = 11008
    num_attention_heads: int = 32
    num_key_value_heads: int = 32
    vocab_size: int = 32064
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0


@dataclass
class LlavaCfg:
    vision: ClipCfg = field(default_factory=ClipCfg)
    text: LlamaCfg = field(default_factory=LlamaCfg)
    image_token_index: int = 32000
    vision_feature_layer: int = -2
    vision_feature_select_strategy: str = "default"


def llava_15_7b() -> LlavaCfg:
    """config C3 shapes."""
    return LlavaCfg()


def tiny_cfg() -> LlavaCfg:
    return LlavaCfg(vision=ClipCfg(hidden_size=64, num_hidden_layers=3, intermediate_size=128,
                                   num_attention_heads=4, image_size=42, patch_size=14),
                    text=LlamaCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=128,
                                  num_attention_heads=4, num_key_value_heads=2, vocab_size=320),
                    image_token_index=300)


# ---------------------------------------------------------------------------
# weights — the reference's post-`sanitize` names (vision.py:196-221: conv weight [O,kH,kW,C])
# ---------------------------------------------------------------------------
def weight_shapes(cfg: LlavaCfg) -> Dict[str, Tuple[int, ...]]:
    v, t = cfg.vision, cfg.text
    E, I = v.hidden_size, v.intermediate_size
    s: Dict[str, Tuple[int, ...]] = {}
    p = "vision_tower.vision_model."
    s[p + "embeddings.class_embedding"] = (E,)
    s[p + "embeddings.patch_embedding.weight"] = (E, v.patch_size, v.patch_size, v.num_channels)
    s[p + "embeddings.position_embedding.weight"] = (v.num_patches + 1, E)
    for n in ("pre_layrnorm", "post_layernorm"):
        s[p + n + ".weight"] = (E,)
        s[p + n + ".bias"] = (E,)
    for i in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            s[q + n + ".weight"] = (E,)
            s[q + n + ".bias"] = (E,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[q + f"self_attn.{n}.weight"] = (E, E)
            s[q + f"self_attn.{n}.bias"] = (E,)
        s[q + "mlp.fc1.weight"], s[q + "mlp.fc1.bias"] = (I, E), (I,)
        s[q + "mlp.fc2.weight"], s[q + "mlp.fc2.bias"] = (E, I), (E,)
    H = t.hidden_size
    s["multi_modal_projector.linear_1.weight"], s["multi_modal_projector.linear_1.bias"] = (H, E), (H,)
    s["multi_modal_projector.linear_2.weight"], s["multi_modal_projector.linear_2.bias"] = (H, H), (H,)
    hd = H // t.num_attention_heads
    kvd = t.num_key_value_heads * hd
    s["language_model.model.embed_tokens.weight"] = (t.vocab_size, H)
    for i in range(t.num_hidden_layers):
        q = f"language_model.model.layers.{i}."
        s[q + "input_layernorm.weight"] = (H,)
        s[q + "post_attention_layernorm.weight"] = (H,)
        s[q + "self_attn.q_proj.weight"] = (H, H)
        s[q + "self_attn.k_proj.weight"] = (kvd, H)
        s[q + "self_attn.v_proj.weight"] = (kvd, H)
        s[q + "self_attn.o_proj.weight"] = (H, H)
        s[q + "mlp.gate_proj.weight"] = (t.intermediate_size, H)
        s[q + "mlp.up_proj.weight"] = (t.intermediate_size, H)
        s[q + "mlp.down_proj.weight"] = (H, t.intermediate_size)
    s["language_model.model.norm.weight"] = (H,)
    s["language_model.lm_head.weight"] = (t.vocab_size, H)
    return s


def init_weights(cfg: LlavaCfg, seed: int = 0, std: float = 0.02, norm_jitter: float = 0.05):
    """bf16-representable fp32 tensors, one generator per tensor (order independent)."""
    W = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        is_norm_w = name.endswith("norm.weight") or ("layer_norm" in name and name.endswith(".weight")) \
            or name.endswith("layrnorm.weight") or name.endswith("layernorm.weight")
        if is_norm_w:
            x = 1.0 + norm_jitter * torch.randn(shape, generator=g)
        else:
            x = std * torch.randn(shape, generator=g)
        W[name] = x.to(torch.bfloat16).to(torch.float32)
    return W


# ---------------------------------------------------------------------------
# vision tower (vision.py:108-185)
# ---------------------------------------------------------------------------
def clip_forward(cfg: LlavaCfg, W, pixel_values_nhwc: torch.Tensor, R: Rounder) -> List[torch.Tensor]:
    """pixel_values (B, H, W, C) -> encoder_states [after pre-LN, after each layer] (B, P+1, E)."""
    v = cfg.vision
    p = "vision_tower.vision_model."
    x = pixel_values_nhwc.to(torch.float32)
    B, Hh, Ww, C = x.shape
    ps = v.patch_size
    gh, gw = Hh // ps, Ww // ps
    # Conv2d(kernel == stride, no bias) == a Linear over the (kH, kW, C)-ordered patch
    patches = x.reshape(B, gh, ps, gw, ps, C).permute(0, 1, 3, 2, 4, 5).reshape(B, gh * gw, ps * ps * C)
    wconv = W[p + "embeddings.patch_embedding.weight"].reshape(v.hidden_size, -1)
    emb = S.linear(R, patches, wconv)
    cls = W[p + "embeddings.class_embedding"].reshape(1, 1, -1).expand(B, 1, -1)
    emb = torch.cat([cls, emb], dim=1)
    emb = R.r(emb + W[p + "embeddings.position_embedding.weight"][None])
    h = S.layer_norm(R, emb, W[p + "pre_layrnorm.weight"], W[p + "pre_layrnorm.bias"], 1e-5)
    # (nn.LayerNorm(hidden) of pre_layrnorm / post_layernorm uses the default eps 1e-5; the encoder
    #  layers use config.layer_norm_eps — vision.py:84-88,149-151)
    states = [h]
    nh = v.num_attention_heads
    hd = v.hidden_size // nh
    for i in range(v.num_hidden_layers):
        q = p + f"encoder.layers.{i}."
        y = S.layer_norm(R, h, W[q + "layer_norm1.weight"], W[q + "layer_norm1.bias"], v.layer_norm_eps)
        qq = S.linear(R, y, W[q + "self_attn.q_proj.weight"], W[q + "self_attn.q_proj.bias"])
        kk = S.linear(R, y, W[q + "self_attn.k_proj.weight"], W[q + "self_attn.k_proj.bias"])
        vv = S.linear(R, y, W[q + "self_attn.v_proj.weight"], W[q + "self_attn.v_proj.bias"])
        L = y.shape[1]
        qq = qq.reshape(B, L, nh, hd).transpose(1, 2)
        kk = kk.reshape(B, L, nh, hd).transpose(1, 2)
        vv = vv.reshape(B, L, nh, hd).transpose(1, 2)
        o = S.sdpa(R, qq, kk, vv, hd ** -0.5, causal=False)
        o = o.transpose(1, 2).reshape(B, L, v.hidden_size)
        o = S.linear(R, o, W[q + "self_attn.out_proj.weight"], W[q + "self_attn.out_proj.bias"])
        h = R.r(h + o)
        y = S.layer_norm(R, h, W[q + "layer_norm2.weight"], W[q + "layer_norm2.bias"], v.layer_norm_eps)
        y = S.linear(R, y, W[q + "mlp.fc1.weight"], W[q + "mlp.fc1.bias"])
        y = S.gelu_fast(R, y)
        y = S.linear(R, y, W[q + "mlp.fc2.weight"], W[q + "mlp.fc2.bias"])
        h = R.r(h + y)
        states.append(h)
    return states


def image_features(cfg: LlavaCfg, W, pixel_values_nhwc, R: Rounder) -> torch.Tensor:
    """llava.py:59-83: select the feature layer, drop CLS, project."""
    states = clip_forward(cfg, W, pixel_values_nhwc, R)
    sel = states[cfg.vision_feature_layer]
    if cfg.vision_feature_select_strategy == "default":
        sel = sel[:, 1:]
    y = S.linear(R, sel, W["multi_modal_projector.linear_1.weight"], W["multi_modal_projector.linear_1.bias"])
    y = S.gelu_exact(R, y)
    return S.linear(R, y, W["multi_modal_projector.linear_2.weight"], W["multi_modal_projector.linear_2.bias"])


def merge_positions(cfg: LlavaCfg, input_ids) -> List[int]:
    ids = np.asarray(input_ids)
    return np.where(ids == cfg.image_token_index)[1].tolist()


def merge_input_ids_with_image_features(cfg: LlavaCfg, image_feats, inputs_embeds, input_ids):
    """llava.py:90-116 (batch 1): the k-th <image> position takes the k-th feature row; more
    feature rows than positions is an error, fewer is numpy's own broadcasting error."""
    pos = merge_positions(cfg, input_ids)
    flat = image_feats.reshape(-1, image_feats.shape[-1])
    if flat.shape[0] > len(pos):
        raise ValueError("Llava model supports only one image per input. Please check your input_ids and pixel_values.")
    out = inputs_embeds.clone()
    out[:, pos, :] = flat.to(out.dtype)
    return out


# ---------------------------------------------------------------------------
# language model: Llama == Qwen2 layers with zero q/k/v bias and 1-D rotary positions
# ---------------------------------------------------------------------------
def _as_qwen(cfg: LlavaCfg, W):
    t = cfg.text
    hd = t.hidden_size // t.num_attention_heads
    tc = Q.TextCfg(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                   intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                   num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                   rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta,
                   mrope_section=(hd // 2, 0, 0), tie_word_embeddings=False)
    qc = Q.Cfg(text=tc, vision=Q.VisionCfg(), image_token_id=-1, video_token_id=-2,
               vision_start_token_id=-3, vision_end_token_id=-4)
    W2 = dict(W)
    kvd = t.num_key_value_heads * hd
    for i in range(t.num_hidden_layers):
        q = f"language_model.model.layers.{i}.self_attn."
        W2[q + "q_proj.bias"] = torch.zeros(t.hidden_size)
        W2[q + "k_proj.bias"] = torch.zeros(kvd)
        W2[q + "v_proj.bias"] = torch.zeros(kvd)
    return qc, W2


def _positions(offset: int, L: int) -> np.ndarray:
    p = np.arange(offset, offset + L)[None, :]
    return np.broadcast_to(p[None], (3, 1, L)).copy()


def get_input_embeddings(cfg: LlavaCfg, W, input_ids, pixel_values_nhwc, R: Rounder,
                         vision_dtype: str = "f32"):
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    embeds = W["language_model.model.embed_tokens.weight"][ids]
    if pixel_values_nhwc is None:
        return embeds, None
    feats = image_features(cfg, W, pixel_values_nhwc, Rounder(vision_dtype))
    feats = R.r(feats)  # astype(inputs_embeds.dtype), llava.py:101-104
    return merge_input_ids_with_image_features(cfg, feats, embeds, input_ids), feats


def greedy_generate(cfg: LlavaCfg, W, input_ids, pixel_values_nhwc, max_tokens: int,
                    dtype: str = "bf16", vision_dtype: str = "f32"):
    """prefill + greedy decode; returns tokens, per-step logits and the prefill pieces."""
    R = Rounder(dtype)
    qc, W2 = _as_qwen(cfg, W)
    embeds, feats = get_input_embeddings(cfg, W, input_ids, pixel_values_nhwc, R, vision_dtype)
    T = embeds.shape[1]
    cache = [Q.OracleKVCache() for _ in range(cfg.text.num_hidden_layers)]
    hidden = Q.lm_layers_forward(qc, W2, embeds, _positions(0, T), cache, R)
    logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
    out_logits, toks = [logits], []
    for n in range(max_tokens):
        lp = Q.logprobs_from_logits(R, logits)
        y = S.argmax_lowest(lp)
        toks.append(int(y[0]))
        if n == max_tokens - 1:
            break
        e = W["language_model.model.embed_tokens.weight"][y][:, None, :]
        hidden = Q.lm_layers_forward(qc, W2, e, _positions(cache[0].offset, 1), cache, R)
        logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
        out_logits.append(logits)
    return {"tokens": toks, "logits": out_logits, "image_features": feats, "inputs_embeds": embeds}


def synthetic_request(cfg: LlavaCfg, n_text: int = 8, seed: int = 0):
    """ids = [text..., <image> x num_patches, text...], pixel_values NHWC float32."""
    rng = np.random.default_rng(seed)
    v = cfg.vision
    lo = rng.integers(3, min(cfg.image_token_index, cfg.text.vocab_size) - 1, size=n_text).tolist()
    ids = lo[: n_text // 2] + [cfg.image_token_index] * v.num_patches + lo[n_text // 2:]
    pv = rng.standard_normal((1, v.image_size, v.image_size, v.num_channels)).astype(np.float32)
    return {"input_ids": np.asarray([ids]), "pixel_values": torch.from_numpy(pv)}
