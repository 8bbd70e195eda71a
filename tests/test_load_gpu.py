"""`load()` round trip (reference utils.py:736-987,1065-1119; vision.py:292-310): a tiny Qwen2-VL
checkpoint is WRITTEN as safetensors + config.json in the HF layout (`visual.*`, `model.*`, conv weight
[O,C,T,H,W]) and in the mlx-community layout (`vision_tower.*`, `language_model.*`, conv weight
[O,T,H,W,C], two shards), read back through `load()`, and must reproduce the oracle run on the same
weights; all engine weights are views of one packed arena."""
import json
import os

import numpy as np
import pytest
import torch

from _util import cmp_noise
from test_engine_gpu import _mk_cfg, _token_ok

pytestmark = pytest.mark.gpu


def _hf_config(c):
    t, v = c.text, c.vision
    return {"model_type": "qwen2_vl", "hidden_size": t.hidden_size, "num_hidden_layers": t.num_hidden_layers,
            "intermediate_size": t.intermediate_size, "num_attention_heads": t.num_attention_heads,
            "num_key_value_heads": t.num_key_value_heads, "rms_norm_eps": t.rms_norm_eps, "vocab_size": t.vocab_size,
            "rope_theta": t.rope_theta, "rope_scaling": {"type": "mrope", "mrope_section": list(t.mrope_section)},
            "tie_word_embeddings": t.tie_word_embeddings, "image_token_id": c.image_token_id,
            "video_token_id": c.video_token_id, "vision_start_token_id": c.vision_start_token_id,
            "vision_config": {"depth": v.depth, "embed_dim": v.embed_dim, "hidden_size": v.hidden_size,
                              "num_heads": v.num_heads, "patch_size": v.patch_size, "mlp_ratio": v.mlp_ratio,
                              "spatial_merge_size": v.spatial_merge_size, "temporal_patch_size": v.temporal_patch_size}}


@pytest.mark.parametrize("layout", ["hf", "mlx"])
def test_load_round_trip(tmp_path, layout):
    from safetensors.torch import save_file
    from mlx_vlm_b200 import load
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.qwen2_vl.processing_qwen2_vl import SyntheticProcessor
    from oracle import qwen2vl as O
    c = _mk_cfg("tiny")
    W = O.init_weights(c, 3, norm_jitter=0.05)
    v = c.vision
    conv = W["vision_tower.patch_embed.proj.weight"].reshape(v.embed_dim, 3, v.temporal_patch_size, v.patch_size,
                                                             v.patch_size)
    tensors = {}
    for k, x in W.items():
        x = x.to(torch.bfloat16)
        if layout == "hf":      # names of the HF checkpoint: visual.*, model.*, lm_head.*
            name = k.replace("vision_tower.", "visual.").replace("language_model.", "")
            if "patch_embed.proj.weight" in k:
                x = conv.to(torch.bfloat16)                          # [O, C, T, H, W]
        else:                   # mlx-community conversion: reference names, channels-last conv
            name = k
            if "patch_embed.proj.weight" in k:
                x = conv.permute(0, 2, 3, 4, 1).to(torch.bfloat16)   # [O, T, H, W, C]
        tensors[name] = x.contiguous()
    d = str(tmp_path)
    names = sorted(tensors)
    if layout == "mlx":         # two shards, like a converted checkpoint
        half = len(names) // 2
        save_file({n: tensors[n] for n in names[:half]}, os.path.join(d, "model-00001-of-00002.safetensors"))
        save_file({n: tensors[n] for n in names[half:]}, os.path.join(d, "model-00002-of-00002.safetensors"))
    else:
        save_file(tensors, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(_hf_config(c), f)
    with open(os.path.join(d, "generation_config.json"), "w") as f:
        json.dump({"eos_token_id": [7, 9]}, f)
    with pytest.raises(NotImplementedError):
        load(d, revision="main")
    from mlx_vlm_b200.models.qwen2_vl.config import ModelConfig
    proc = SyntheticProcessor(ModelConfig.from_dict(_hf_config(c)), n_text_tokens=12, seed=0)
    model, processor = load(d, processor=proc, device="cuda:0")
    assert processor is proc and list(model.config.eos_token_id) == [7, 9]
    eng = model.engine
    flat = model.packed_weights
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 2
    assert all(lo <= t.data_ptr() < hi for t in eng.weights.values()), "every weight is a view of the arena"
    req = O.synthetic_request(c, 12, image_hw=(56, 56), seed=1)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    n = 5
    ref = O.greedy_generate(c, W, ids, pv, grid, n)
    toks = ref["tokens"][0].tolist()
    ex = O.greedy_generate(c, W, ids, pv, grid, n, dtype="f32", force_tokens=toks[:])
    pvd = torch.from_numpy(pv).cuda()
    for i, (tok, lp) in enumerate(generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid)):
        assert _token_ok(tok, ref["logprobs"][i][0]), f"{layout}: token {i}: {tok} vs {toks[i]}"
        if tok != toks[i]:
            break
        cmp_noise(lp, ref["logprobs"][i][0], ex["logprobs"][i][0], f"load({layout}) logprobs step {i}")


@pytest.mark.parametrize("family", ["llava", "idefics2"])
def test_load_round_trip_other_families(tmp_path, family):
    """`load()` dispatches on `model_type` (utils.py:588-635) to the LLaVA-1.5 / Idefics2 packages: a tiny checkpoint
    written as safetensors (PyTorch conv layout [O,C,kH,kW], plus the keys the reference drops: `position_ids`,
    `rotary_emb.inv_freq`; Idefics2 with the HF `model.` / `text_model.` / `lm_head.` prefixes) is read back through
    `sanitize` + `load_weights` and must reproduce the oracle's greedy tokens / step-0 logprobs."""
    from safetensors.torch import save_file
    from mlx_vlm_b200 import load
    from mlx_vlm_b200.generate import generate_step
    from oracle.mlx_semantics import Rounder
    d = str(tmp_path)
    tensors = {}
    if family == "llava":
        from oracle import llava as OM
        c = OM.LlavaCfg(vision=OM.ClipCfg(hidden_size=64, num_hidden_layers=3, intermediate_size=128, num_attention_heads=4,
                                          image_size=42, patch_size=14),
                        text=OM.LlamaCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                                         num_key_value_heads=2, vocab_size=320), image_token_index=300)
        W = OM.init_weights(c, 5)
        for k, x in W.items():
            if "patch_embedding.weight" in k:
                x = x.permute(0, 3, 1, 2)                                   # [O,kH,kW,C] -> PyTorch [O,C,kH,kW]
            tensors[k] = x.to(torch.bfloat16).contiguous()
        tensors["vision_tower.vision_model.embeddings.position_ids"] = torch.arange(10).to(torch.bfloat16)
        tensors["language_model.model.layers.0.self_attn.rotary_emb.inv_freq"] = torch.ones(4, dtype=torch.bfloat16)
        v, t = c.vision, c.text
        cfg = {"model_type": "llava", "image_token_index": c.image_token_index, "vision_feature_layer": c.vision_feature_layer,
               "vision_feature_select_strategy": c.vision_feature_select_strategy, "vocab_size": t.vocab_size,
               "text_config": {"model_type": "llama", "hidden_size": t.hidden_size, "num_hidden_layers": t.num_hidden_layers,
                               "intermediate_size": t.intermediate_size, "num_attention_heads": t.num_attention_heads,
                               "num_key_value_heads": t.num_key_value_heads, "vocab_size": t.vocab_size,
                               "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta},
               "vision_config": {"model_type": "clip_vision_model", "num_hidden_layers": v.num_hidden_layers,
                                 "hidden_size": v.hidden_size, "intermediate_size": v.intermediate_size,
                                 "num_attention_heads": v.num_attention_heads, "image_size": v.image_size,
                                 "patch_size": v.patch_size, "layer_norm_eps": v.layer_norm_eps}}
        req = OM.synthetic_request(c, n_text=8, seed=2)
        ids = req["input_ids"]
        pv = req["pixel_values"].permute(0, 3, 1, 2).contiguous().cuda()
        ref = OM.greedy_generate(c, W, ids, req["pixel_values"], 3)
        extra = {}
    else:
        from oracle import idefics2 as OM
        c = OM.Idefics2Cfg(
            vision=OM.SiglipCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=96, num_attention_heads=4,
                                image_size=70, patch_size=14),
            text=OM.MistralCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                               num_key_value_heads=2, vocab_size=320),
            perceiver=OM.PerceiverCfg(num_key_value_heads=2, resampler_depth=2, resampler_head_dim=16, resampler_n_heads=4,
                                      resampler_n_latents=6), image_token_index=300)
        W = OM.init_weights(c, 5)
        for k, x in W.items():
            if "patch_embedding.weight" in k:
                x = x.permute(0, 3, 1, 2)
            if k.startswith("language_model.lm_head"):
                name = "lm_head." + k.split("lm_head.", 1)[1]
            elif k.startswith("language_model."):
                name = "model.text_model." + k[len("language_model."):]
            else:
                name = "model." + k
            tensors[name] = x.to(torch.bfloat16).contiguous()
        v, t, p = c.vision, c.text, c.perceiver
        cfg = {"model_type": "idefics2", "image_token_id": c.image_token_index, "vocab_size": t.vocab_size,
               "text_config": {"model_type": "mistral", "hidden_size": t.hidden_size, "num_hidden_layers": t.num_hidden_layers,
                               "intermediate_size": t.intermediate_size, "num_attention_heads": t.num_attention_heads,
                               "num_key_value_heads": t.num_key_value_heads, "vocab_size": t.vocab_size,
                               "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta},
               "vision_config": {"hidden_size": v.hidden_size, "num_hidden_layers": v.num_hidden_layers,
                                 "intermediate_size": v.intermediate_size, "num_attention_heads": v.num_attention_heads,
                                 "image_size": v.image_size, "patch_size": v.patch_size, "layer_norm_eps": v.layer_norm_eps},
               "perceiver_config": {"num_key_value_heads": p.num_key_value_heads, "resampler_depth": p.resampler_depth,
                                    "resampler_head_dim": p.resampler_head_dim, "resampler_n_heads": p.resampler_n_heads,
                                    "resampler_n_latents": p.resampler_n_latents}}
        req = OM.synthetic_request(c, n_images=1, n_text=8, seed=2)
        ids = req["input_ids"]
        pv = torch.from_numpy(req["pixel_values"]).cuda()
        ref = OM.greedy_generate(c, W, ids, req["pixel_values"], None, 3)
        extra = {}
    save_file(tensors, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    import types
    proc = types.SimpleNamespace(tokenizer=types.SimpleNamespace(stopping_criteria=None))
    model, processor = load(d, processor=proc, device="cuda:0")
    assert type(model).__module__.endswith(f"models.{family}.{family}") and processor is proc
    for i, (tok, lp) in enumerate(generate_step(ids, model, pv, None, max_tokens=3, **extra)):
        lp_ref = OM.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        assert _token_ok(tok, lp_ref), f"{family}: token {i}: {tok} vs {ref['tokens'][i]}"
        if tok != ref["tokens"][i]:
            break
    assert model.engine.device_error() == 0
