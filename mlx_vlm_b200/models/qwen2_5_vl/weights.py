"""Seeded random weights at the Qwen2.5-VL shapes under the reference's names (benchmarks / tests)."""
from __future__ import annotations

import torch


def weight_shapes(config):
    v, t = config.vision_config, config.text_config
    E, I, H = v.hidden_size, v.intermediate_size, t.hidden_size
    s = {"vision_tower.patch_embed.proj.weight": (E, v.in_channels, v.temporal_patch_size, v.patch_size, v.patch_size)}
    for i in range(v.depth):
        p = f"vision_tower.blocks.{i}."
        s[p + "norm1.weight"], s[p + "norm2.weight"] = (E,), (E,)
        s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"] = (3 * E, E), (3 * E,)
        s[p + "attn.proj.weight"], s[p + "attn.proj.bias"] = (E, E), (E,)
        for n, shp in (("gate_proj", (I, E)), ("up_proj", (I, E)), ("down_proj", (E, I))):
            s[p + f"mlp.{n}.weight"], s[p + f"mlp.{n}.bias"] = shp, (shp[0],)
    m = E * v.spatial_merge_size ** 2
    s["vision_tower.merger.ln_q.weight"] = (E,)
    s["vision_tower.merger.mlp.0.weight"], s["vision_tower.merger.mlp.0.bias"] = (m, m), (m,)
    s["vision_tower.merger.mlp.2.weight"], s["vision_tower.merger.mlp.2.bias"] = (v.out_hidden_size, m), (v.out_hidden_size,)
    hd = H // t.num_attention_heads
    kv = t.num_key_value_heads * hd
    s["language_model.model.embed_tokens.weight"] = (t.vocab_size, H)
    for i in range(t.num_hidden_layers):
        p = f"language_model.model.layers.{i}."
        s[p + "input_layernorm.weight"], s[p + "post_attention_layernorm.weight"] = (H,), (H,)
        for n, rows in (("q", H), ("k", kv), ("v", kv)):
            s[p + f"self_attn.{n}_proj.weight"], s[p + f"self_attn.{n}_proj.bias"] = (rows, H), (rows,)
        s[p + "self_attn.o_proj.weight"] = (H, H)
        s[p + "mlp.gate_proj.weight"], s[p + "mlp.up_proj.weight"] = (t.intermediate_size, H), (t.intermediate_size, H)
        s[p + "mlp.down_proj.weight"] = (H, t.intermediate_size)
    s["language_model.model.norm.weight"] = (H,)
    if not t.tie_word_embeddings:
        s["language_model.lm_head.weight"] = (t.vocab_size, H)
    return s


def random_weights(config, seed=0, std=0.02, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    W = {}
    for name, shape in weight_shapes(config).items():
        if len(shape) == 1 and ("norm" in name or "ln_q" in name):
            W[name] = torch.ones(shape, device=device, dtype=torch.bfloat16)
        else:
            W[name] = (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)
    return W
