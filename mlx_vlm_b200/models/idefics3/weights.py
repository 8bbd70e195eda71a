"""Seeded random weights at the Idefics3 shapes under the reference's names (benchmarks / tests)."""
from __future__ import annotations

import torch


def weight_shapes(config):
    v, t = config.vision_config, config.text_config
    E, I, H = v.hidden_size, v.intermediate_size, t.hidden_size
    s = {}
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
    s["connector.modality_projection.proj.weight"] = (H, E * config.scale_factor ** 2)
    hd = H // t.num_attention_heads
    lkv = t.num_key_value_heads * hd
    s["language_model.embed_tokens.weight"] = (t.vocab_size, H)
    for i in range(t.num_hidden_layers):
        q = f"language_model.layers.{i}."
        s[q + "input_layernorm.weight"], s[q + "post_attention_layernorm.weight"] = (H,), (H,)
        s[q + "self_attn.q_proj.weight"], s[q + "self_attn.o_proj.weight"] = (H, H), (H, H)
        s[q + "self_attn.k_proj.weight"], s[q + "self_attn.v_proj.weight"] = (lkv, H), (lkv, H)
        s[q + "mlp.gate_proj.weight"], s[q + "mlp.up_proj.weight"] = (t.intermediate_size, H), (t.intermediate_size, H)
        s[q + "mlp.down_proj.weight"] = (H, t.intermediate_size)
    s["language_model.norm.weight"] = (H,)
    s["language_model.lm_head.weight"] = (t.vocab_size, H)
    return s


def random_weights(config, seed=0, std=0.02, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    W = {}
    for name, shape in weight_shapes(config).items():
        stem = name.split(".")[-2] if "." in name else name
        if name.endswith(".weight") and "norm" in stem:
            W[name] = torch.ones(shape, device=device, dtype=torch.bfloat16)
        elif name.endswith(".bias") and "norm" in stem:
            W[name] = torch.zeros(shape, device=device, dtype=torch.bfloat16)
        else:
            W[name] = (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)
    return W
