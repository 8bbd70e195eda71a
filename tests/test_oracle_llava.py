"""CPU: the LLaVA-1.5 oracle (oracle/llava.py) is pinned before any kernel is
written for it:
  * merge indexing against the reference's own function source (tests/golden/make_golden.py);
  * the whole wiring (CLIP tower incl. class token / positions / pre-LN / feature layer -2 / CLS
    drop, projector, merge, Llama with rotary + GQA) in fp32 against HuggingFace transformers'
    LlavaForConditionalGeneration with the same weights."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import llava as L
from oracle.mlx_semantics import Rounder

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    GOLD = json.load(f)


def test_llava_merge_matches_reference_source():
    cfg = L.tiny_cfg()
    cfg.image_token_index = 100
    for case in GOLD["llava_merge"]:
        ids = np.asarray(case["input_ids"])
        H, n = case["hidden"], case["n_feats"]
        feats = torch.from_numpy((1000 + np.arange(n * H, dtype=np.float32)).reshape(1, n, H))
        emb = torch.from_numpy(-(np.arange(ids.size * H, dtype=np.float32) + 1).reshape(1, ids.shape[1], H))
        if case["error"] is not None:
            with pytest.raises((ValueError, RuntimeError, IndexError)):
                L.merge_input_ids_with_image_features(cfg, feats, emb, ids)
            continue
        out = L.merge_input_ids_with_image_features(cfg, feats, emb, ids)
        assert np.array_equal(out.numpy(), np.asarray(case["output"], dtype=np.float32)), case["tag"]
        assert L.merge_positions(cfg, ids) == case["positions"]


def _hf_model(c):
    transformers = pytest.importorskip("transformers")
    from transformers import CLIPVisionConfig, LlamaConfig, LlavaConfig, LlavaForConditionalGeneration
    v, t = c.vision, c.text
    vc = CLIPVisionConfig(hidden_size=v.hidden_size, num_hidden_layers=v.num_hidden_layers,
                          intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                          image_size=v.image_size, patch_size=v.patch_size, num_channels=v.num_channels,
                          layer_norm_eps=v.layer_norm_eps, hidden_act="quick_gelu", projection_dim=32)
    tc = LlamaConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                     intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                     num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                     rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta, tie_word_embeddings=False,
                     max_position_embeddings=512, attention_bias=False, mlp_bias=False)
    cfg = LlavaConfig(vision_config=vc, text_config=tc, image_token_index=c.image_token_index,
                      vision_feature_layer=c.vision_feature_layer,
                      vision_feature_select_strategy=c.vision_feature_select_strategy,
                      projector_hidden_act="gelu", image_seq_length=v.num_patches)
    torch.manual_seed(0)
    return LlavaForConditionalGeneration(cfg).eval().float()


def _load_into_hf(m, W):
    sd = m.state_dict()
    new = {}
    for k in sd:
        kk = k
        if k.startswith("model.vision_tower."):
            kk = "vision_tower." + k[len("model.vision_tower."):]
        elif k.startswith("model.multi_modal_projector."):
            kk = "multi_modal_projector." + k[len("model.multi_modal_projector."):]
        elif k.startswith("model.language_model."):
            kk = "language_model.model." + k[len("model.language_model."):]
        elif k.startswith("lm_head."):
            kk = "language_model." + k
        if kk.endswith("position_ids") or kk not in W:
            new[k] = sd[k]
            continue
        x = W[kk]
        if kk.endswith("patch_embedding.weight"):
            x = x.permute(0, 3, 1, 2).contiguous()  # [O,kH,kW,C] (mlx) -> [O,C,kH,kW] (torch)
        assert tuple(x.shape) == tuple(sd[k].shape), (k, x.shape, sd[k].shape)
        new[k] = x.clone()
    missing = [k for k in L.weight_shapes(L.tiny_cfg()) if "post_layernorm" not in k and
               not any(k == (kk.replace("model.vision_tower.", "vision_tower.")
                             .replace("model.multi_modal_projector.", "multi_modal_projector.")
                             .replace("model.language_model.", "language_model.model.")
                             if not kk.startswith("lm_head.") else "language_model." + kk) for kk in sd)]
    assert not missing, missing[:5]
    m.load_state_dict(new)


def test_llava_oracle_f32_matches_hf_transformers():
    c = L.tiny_cfg()
    try:
        m = _hf_model(c)
    except Exception as e:  # config API drift between transformers versions
        pytest.skip(f"cannot build the HF model here: {e}")
    W = L.init_weights(c, seed=3)
    _load_into_hf(m, W)
    req = L.synthetic_request(c, n_text=10, seed=1)
    ids, pv = req["input_ids"], req["pixel_values"]
    out = L.greedy_generate(c, W, ids, pv, 1, dtype="f32", vision_dtype="f32")
    with torch.no_grad():
        hf = m(input_ids=torch.from_numpy(ids), pixel_values=pv.permute(0, 3, 1, 2).contiguous(),
               attention_mask=torch.ones_like(torch.from_numpy(ids)))
    want = hf.logits[0, -1].float()
    got = out["logits"][0][0]
    rel = float((got - want).norm() / want.norm())
    print(f"LLaVA oracle f32 vs HF logits rel_l2={rel:.3e}")
    assert rel <= 2e-5
    # text-only request through the same LM
    tids = np.asarray([[5, 9, 17, 33, 2]])
    o2 = L.greedy_generate(c, W, tids, None, 1, dtype="f32")
    with torch.no_grad():
        h2 = m(input_ids=torch.from_numpy(tids))
    assert float((o2["logits"][0][0] - h2.logits[0, -1]).norm() / h2.logits[0, -1].norm()) <= 2e-5


def test_llava_vision_tower_is_fp32_in_the_reference():
    """The reference feeds float32 pixel_values to bf16 weights (llava.py:61-63, utils.py:2091): the
    tower runs in fp32.  A bf16 tower is a different (measurably noisier) function — quantified here
    so that the round-2 kernel choice is explicit."""
    c = L.tiny_cfg()
    W = L.init_weights(c, seed=1)
    pv = L.synthetic_request(c, seed=2)["pixel_values"]
    f32 = L.image_features(c, W, pv, Rounder("f32"))
    bf = L.image_features(c, W, pv, Rounder("bf16"))
    rel = float((bf - f32).norm() / f32.norm())
    print(f"bf16 tower vs the reference's fp32 tower: rel_l2={rel:.3e}")
    assert 1e-4 < rel < 5e-2
