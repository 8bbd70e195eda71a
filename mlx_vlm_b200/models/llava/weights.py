"""Seeded random weights at the LLaVA shapes under the reference's names (benchmarks / tests)."""
from __future__ import annotations

import torch


def weight_shapes(config):
    v, t = config.vision_config, config.text_config
    E, I = v.hidden_size, v.intermediate_size
    P = (v.image_size // v.patch_size) ** 2
    s = {}
    p = "vision_tower.vision_model."
    s[p + "embeddings.class_embedding"] = (E,)
    s[p + "embeddings.patch_embedding.weight"] = (E, v.patch_size, v.patch_size, v.num_channels)
    s[p + "embeddings.position_embedding.weight"] = (P + 1, E)
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


def random_weights(config, seed=0, std=0.02, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    W = {}
    for name, shape in weight_shapes(config).items():
        is_norm = name.endswith(("norm.weight", "layrnorm.weight", "layernorm.weight")) or \
            ("layer_norm" in name and name.endswith(".weight"))
        if is_norm:
            W[name] = torch.ones(shape, device=device, dtype=torch.bfloat16)
        elif name.endswith(("norm.bias", "layrnorm.bias", "layernorm.bias")) or \
                ("layer_norm" in name and name.endswith(".bias")):
            W[name] = torch.zeros(shape, device=device, dtype=torch.bfloat16)
        else:
            W[name] = (torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std).to(torch.bfloat16)
    return W
