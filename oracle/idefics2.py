"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's Idefics2 generate path:
  models/idefics2/vision.py:123-173  patch embedding (Conv2d 14x14 + bias) and the BUCKETED
                                     fractional position ids (`np.digitize`; integer, bit-exact)
  models/idefics2/vision.py:20-120   SigLIP encoder: N x {LN, MHA with bias, LN, fast-GELU MLP}
  models/idefics2/vision.py:176-215  VisionModel: post-LayerNorm of the last hidden state
  models/idefics2/idefics2.py:36-171 connector: modality MLP (SwiGLU) + Perceiver resampler
                                     (latents attend to concat[context, latents], GQA)
  models/idefics2/idefics2.py:185-262 get_input_embeddings: padding-image removal, pixel mask ->
                                     patch mask, vision -> connector -> masked_scatter merge
  models/idefics2/language.py:16-150 Mistral decoder (nn.RoPE, no biases, untied head)

PARITY STATUS: oracle only — the product kernels for this row are not built yet (round 2).
Pinned (tests/test_oracle_idefics2.py): position ids / patch mask / padding-image removal /
merge against the reference's own source (tests/golden/), the whole wiring in fp32 against
HuggingFace transformers' Idefics2ForConditionalGeneration with the same weights.

Precision: like LLaVA, the reference feeds float32 pixel_values to bf16 weights and even casts the
pooled features back to the pixel dtype (`pooler_output.astype(pixel_values.dtype)`,
idefics2.py:251), so tower AND connector run in fp32; the merge writes into the bf16 embeddings.
The reference's vision encoder is called WITHOUT an attention mask (vision.py:207): padded patches
take part in attention; only their position id stays 0.

Two places where the reference differs from the HuggingFace model it was ported from (the oracle
follows the REFERENCE; the HF cross-check substitutes these two):
  * position buckets: `np.digitize(frac, boundaries, right=True) - 1` (vision.py:160-165) is one
    bucket lower than HF's `torch.bucketize(frac, boundaries, right=True)`; coordinate 0 lands in
    bucket -1, so ids can be NEGATIVE and index the position table from its end (golden vectors
    in tests/golden/ record exactly that);
  * the post-LayerNorm is `nn.LayerNorm(hidden)` with the default eps 1e-5, HF uses layer_norm_eps.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder


@dataclass
class SiglipCfg:
    hidden_size: int = 1152
    num_hidden_layers: int = 27
    intermediate_size: int = 4304
    num_attention_heads: int = 16
    image_size: int = 980
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6


@dataclass
class MistralCfg:
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    intermediate_size: int = 14336
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    vocab_size: int = 32003
    rms_norm_eps: float = 1e-5
    rope_theta: float = 1000000.0


@dataclass
class PerceiverCfg:
    num_key_value_heads: int = 4
    resampler_depth: int = 3
    resampler_head_dim: int = 96
    resampler_n_heads: int = 16
    resampler_n_latents: int = 64


@dataclass
class Idefics2Cfg:
    vision: SiglipCfg = field(default_factory=SiglipCfg)
    text: MistralCfg = field(default_factory=MistralCfg)
    perceiver: PerceiverCfg = field(default_factory=PerceiverCfg)
    image_token_index: int = 32001


def idefics2_8b() -> Idefics2Cfg:
    return Idefics2Cfg()


def tiny_cfg() -> Idefics2Cfg:
    return Idefics2Cfg(vision=SiglipCfg(hidden_size=48, num_hidden_layers=2, intermediate_size=96,
                                        num_attention_heads=4, image_size=70, patch_size=14),
                       text=MistralCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=128,
                                       num_attention_heads=4, num_key_value_heads=2, vocab_size=320),
                       perceiver=PerceiverCfg(num_key_value_heads=2, resampler_depth=2,
                                              resampler_head_dim=16, resampler_n_heads=4,
                                              resampler_n_latents=6),
                       image_token_index=300)


# ---------------------------------------------------------------------------
# integer pieces — bit-exact
# ---------------------------------------------------------------------------
def bucketed_position_ids(patch_mask: np.ndarray, num_patches_per_side: int) -> np.ndarray:
    """vision.py:150-170.  patch_mask (B, ph, pw) bool -> (B, ph*pw) int.  The valid top-left
    nh x nw block of an image is mapped onto the num_patches x num_patches position grid by
    bucketing the fractional coordinates i/nh, j/nw; padded patches keep position 0."""
    B, ph, pw = patch_mask.shape
    boundaries = np.linspace(1 / num_patches_per_side, 1.0, num_patches_per_side, endpoint=False)
    out = np.zeros((B, ph * pw), dtype=int)
    for b in range(B):
        m = np.asarray(patch_mask[b])
        nh, nw = int(m[:, 0].sum()), int(m[0, :].sum())
        fh = np.linspace(0, 1, nh, endpoint=False)
        fw = np.linspace(0, 1, nw, endpoint=False)
        bh = np.digitize(fh, boundaries, right=True) - 1
        bw = np.digitize(fw, boundaries, right=True) - 1
        out[b][m.reshape(-1)] = (bh[:, None] * num_patches_per_side + bw).flatten()
    return out


def real_image_indices(pixel_values_bnchw: np.ndarray) -> List[int]:
    """idefics2.py:204-210: an all-zero image is padding and is dropped."""
    pv = np.asarray(pixel_values_bnchw)
    B, N = pv.shape[:2]
    flat = pv.reshape(B * N, -1)
    return np.where((flat == 0.0).sum(axis=1) != flat.shape[1])[0].tolist()


def patch_attention_mask(pixel_attention_mask: np.ndarray, patch_size: int) -> np.ndarray:
    """idefics2.py:226-243: a patch is valid iff any of its pixels is."""
    m = np.asarray(pixel_attention_mask)
    B, H, W = m.shape
    ph, pw = H // patch_size, W // patch_size
    m = m[:, :ph * patch_size, :pw * patch_size].reshape(B, ph, patch_size, pw, patch_size)
    return m.transpose(0, 1, 3, 2, 4).sum(axis=(-1, -2)) > 0


def merge(cfg: Idefics2Cfg, image_features: torch.Tensor, inputs_embeds: torch.Tensor, input_ids):
    """idefics2.py:263-280 + masked_scatter :15-33: the flattened features fill, in order, the
    flattened positions of the <image> rows; element counts must match exactly."""
    ids = np.asarray(input_ids)
    mask = torch.from_numpy(ids == cfg.image_token_index)[..., None].expand_as(inputs_embeds)
    if int(mask.sum()) != image_features.numel():
        raise ValueError(f"Image features and image tokens do not match: tokens: {int((ids == cfg.image_token_index).sum())}, "
                         f"features {image_features.shape[0]}")
    out = inputs_embeds.clone().reshape(-1)
    out[mask.reshape(-1)] = image_features.reshape(-1).to(out.dtype)
    return out.reshape(inputs_embeds.shape)


# ---------------------------------------------------------------------------
# weights (reference attribute names)
# ---------------------------------------------------------------------------
def weight_shapes(cfg: Idefics2Cfg) -> Dict[str, Tuple[int, ...]]:
    v, t, pc = cfg.vision, cfg.text, cfg.perceiver
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
    c = "connector."
    s[c + "modality_projection.gate_proj.weight"] = (t.intermediate_size, E)
    s[c + "modality_projection.up_proj.weight"] = (t.intermediate_size, E)
    s[c + "modality_projection.down_proj.weight"] = (H, t.intermediate_size)
    r = c + "perceiver_resampler."
    s[r + "latents"] = (pc.resampler_n_latents, H)
    qd, kvd = pc.resampler_n_heads * pc.resampler_head_dim, pc.num_key_value_heads * pc.resampler_head_dim
    for i in range(pc.resampler_depth):
        q = r + f"layers.{i}."
        for n in ("input_latents_norm", "input_context_norm", "post_attention_layernorm"):
            s[q + n + ".weight"] = (H,)
        s[q + "self_attn.q_proj.weight"] = (qd, H)
        s[q + "self_attn.k_proj.weight"] = (kvd, H)
        s[q + "self_attn.v_proj.weight"] = (kvd, H)
        s[q + "self_attn.o_proj.weight"] = (H, qd)
        s[q + "mlp.gate_proj.weight"] = (4 * H, H)
        s[q + "mlp.up_proj.weight"] = (4 * H, H)
        s[q + "mlp.down_proj.weight"] = (H, 4 * H)
    s[r + "norm.weight"] = (H,)
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


def init_weights(cfg: Idefics2Cfg, seed: int = 0, std: float = 0.02, norm_jitter: float = 0.05):
    W = {}
    for idx, (name, shape) in enumerate(weight_shapes(cfg).items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        norm_w = name.endswith(".weight") and ("norm" in name.split(".")[-2])
        if norm_w:
            x = 1.0 + norm_jitter * torch.randn(shape, generator=g)
        elif name.endswith("latents"):
            x = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            x = std * torch.randn(shape, generator=g)
        W[name] = x.to(torch.bfloat16).to(torch.float32)
    return W


# ---------------------------------------------------------------------------
# vision tower + connector
# ---------------------------------------------------------------------------
def vision_forward(cfg: Idefics2Cfg, W, pixel_values_nhwc: torch.Tensor, patch_mask: np.ndarray, R: Rounder,
                   position_ids: Optional[np.ndarray] = None, post_ln_eps: float = 1e-5):
    """(n_img, H, W, C) -> post-LayerNormed last hidden state (n_img, ph*pw, E).
    `position_ids` / `post_ln_eps` exist only so that the HF cross-check can substitute HF's
    bucketing and epsilon (the reference differs from HF in both, see the module docstring)."""
    v = cfg.vision
    p = "vision_model."
    x = pixel_values_nhwc.to(torch.float32)
    B, Hh, Ww, C = x.shape
    ps = v.patch_size
    gh, gw = Hh // ps, Ww // ps
    patches = x.reshape(B, gh, ps, gw, ps, C).permute(0, 1, 3, 2, 4, 5).reshape(B, gh * gw, ps * ps * C)
    emb = S.linear(R, patches, W[p + "embeddings.patch_embedding.weight"].reshape(v.hidden_size, -1),
                   W[p + "embeddings.patch_embedding.bias"])
    pos = bucketed_position_ids(patch_mask, v.image_size // ps) if position_ids is None else position_ids
    h = R.r(emb + W[p + "embeddings.position_embedding.weight"][torch.from_numpy(pos)])
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
        y = S.gelu_fast(R, S.linear(R, y, W[q + "mlp.fc1.weight"], W[q + "mlp.fc1.bias"]))
        y = S.linear(R, y, W[q + "mlp.fc2.weight"], W[q + "mlp.fc2.bias"])
        h = R.r(h + y)
    # nn.LayerNorm(hidden_size) -> default eps 1e-5 (vision.py:184; HF uses layer_norm_eps here)
    return S.layer_norm(R, h, W[p + "post_layernorm.weight"], W[p + "post_layernorm.bias"], post_ln_eps)


def _swiglu_mlp(R, x, W, prefix):
    g = S.linear(R, x, W[prefix + "gate_proj.weight"])
    u = S.linear(R, x, W[prefix + "up_proj.weight"])
    return S.linear(R, S.swiglu(R, g, u), W[prefix + "down_proj.weight"])


def connector_forward(cfg: Idefics2Cfg, W, feats: torch.Tensor, R: Rounder) -> torch.Tensor:
    """(n_img, P, E) -> (n_img, n_latents, H)."""
    pc, t = cfg.perceiver, cfg.text
    x = _swiglu_mlp(R, feats, W, "connector.modality_projection.")
    r = "connector.perceiver_resampler."
    B = x.shape[0]
    h = W[r + "latents"][None].expand(B, -1, -1)
    nh, nkv, hd = pc.resampler_n_heads, pc.num_key_value_heads, pc.resampler_head_dim
    for i in range(pc.resampler_depth):
        q = r + f"layers.{i}."
        lat = S.rms_norm(R, h, W[q + "input_latents_norm.weight"], t.rms_norm_eps)
        ctx = S.rms_norm(R, x, W[q + "input_context_norm.weight"], t.rms_norm_eps)
        kvin = torch.cat([ctx, lat], dim=1)
        L, Skv = lat.shape[1], kvin.shape[1]
        qq = S.linear(R, lat, W[q + "self_attn.q_proj.weight"]).reshape(B, L, nh, hd).transpose(1, 2)
        kk = S.linear(R, kvin, W[q + "self_attn.k_proj.weight"]).reshape(B, Skv, nkv, hd).transpose(1, 2)
        vv = S.linear(R, kvin, W[q + "self_attn.v_proj.weight"]).reshape(B, Skv, nkv, hd).transpose(1, 2)
        o = S.sdpa(R, qq, kk, vv, hd ** -0.5, causal=False).transpose(1, 2).reshape(B, L, nh * hd)
        o = S.linear(R, o, W[q + "self_attn.o_proj.weight"])
        h = R.r(h + o)
        y = S.rms_norm(R, h, W[q + "post_attention_layernorm.weight"], t.rms_norm_eps)
        h = R.r(h + _swiglu_mlp(R, y, W, q + "mlp."))
    return S.rms_norm(R, h, W[r + "norm.weight"], t.rms_norm_eps)


def image_features(cfg: Idefics2Cfg, W, pixel_values_bnchw: np.ndarray, pixel_attention_mask, R: Rounder,
                   position_ids: Optional[np.ndarray] = None, post_ln_eps: float = 1e-5):
    pv = np.asarray(pixel_values_bnchw, dtype=np.float32)
    B, N, C, Hh, Ww = pv.shape
    keep = real_image_indices(pv)
    pv = pv.reshape(B * N, C, Hh, Ww)[keep]
    if pixel_attention_mask is None:
        pam = np.ones((pv.shape[0], Hh, Ww), dtype=bool)
    else:
        pam = np.asarray(pixel_attention_mask).reshape(B * N, Hh, Ww)[keep]
    pmask = patch_attention_mask(pam, cfg.vision.patch_size)
    x = torch.from_numpy(pv).permute(0, 2, 3, 1)
    feats = vision_forward(cfg, W, x, pmask, R, position_ids, post_ln_eps)
    return connector_forward(cfg, W, feats, R)


# ---------------------------------------------------------------------------
# language model: Mistral == Qwen2 layers with zero q/k/v bias, 1-D rotary positions
# ---------------------------------------------------------------------------
def _as_qwen(cfg: Idefics2Cfg, W):
    t = cfg.text
    hd = t.hidden_size // t.num_attention_heads
    tc = Q.TextCfg(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                   intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                   num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                   rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta,
                   mrope_section=(hd // 2, 0, 0), tie_word_embeddings=False)
    qc = Q.Cfg(text=tc, vision=Q.VisionCfg(), image_token_id=-1, video_token_id=-2,
               vision_start_token_id=-3, vision_end_token_id=-4)
    W2 = {}
    for k, x in W.items():
        if k.startswith("language_model.") and not k.startswith("language_model.lm_head"):
            W2["language_model.model." + k[len("language_model."):]] = x
        else:
            W2[k] = x
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


def greedy_generate(cfg: Idefics2Cfg, W, input_ids, pixel_values_bnchw, pixel_attention_mask,
                    max_tokens: int, dtype: str = "bf16", vision_dtype: str = "f32",
                    position_ids: Optional[np.ndarray] = None, post_ln_eps: float = 1e-5):
    R = Rounder(dtype)
    qc, W2 = _as_qwen(cfg, W)
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    embeds = W["language_model.embed_tokens.weight"][ids]
    feats = None
    if pixel_values_bnchw is not None:
        feats = R.r(image_features(cfg, W, pixel_values_bnchw, pixel_attention_mask, Rounder(vision_dtype),
                                   position_ids, post_ln_eps))
        embeds = merge(cfg, feats, embeds, input_ids)
    T = embeds.shape[1]
    cache = [Q.OracleKVCache() for _ in range(cfg.text.num_hidden_layers)]
    hidden = Q.lm_layers_forward(qc, W2, embeds, _positions(0, T), cache, R)
    logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
    out_logits, toks = [logits], []
    for n in range(max_tokens):
        y = S.argmax_lowest(Q.logprobs_from_logits(R, logits))
        toks.append(int(y[0]))
        if n == max_tokens - 1:
            break
        e = W["language_model.embed_tokens.weight"][y][:, None, :]
        hidden = Q.lm_layers_forward(qc, W2, e, _positions(cache[0].offset, 1), cache, R)
        logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
        out_logits.append(logits)
    return {"tokens": toks, "logits": out_logits, "image_features": feats, "inputs_embeds": embeds}


def synthetic_request(cfg: Idefics2Cfg, n_images: int = 2, n_text: int = 8, seed: int = 0):
    rng = np.random.default_rng(seed)
    v = cfg.vision
    n_lat = cfg.perceiver.resampler_n_latents
    text = rng.integers(3, cfg.image_token_index - 1, size=n_text).tolist()
    ids = text[: n_text // 2] + [cfg.image_token_index] * (n_lat * n_images) + text[n_text // 2:]
    pv = rng.standard_normal((1, n_images, v.num_channels, v.image_size, v.image_size)).astype(np.float32)
    return {"input_ids": np.asarray([ids]), "pixel_values": pv}
