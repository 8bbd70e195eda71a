"""CPU: the Idefics2 oracle (oracle/idefics2.py) pinned before any kernel is
written for it: the integer logic (bucketed position ids incl. the reference's negative buckets,
pixel mask -> patch mask, padding-image removal, masked_scatter merge) against the reference's own
source (tests/golden/make_golden.py), the wiring in fp32 against HuggingFace transformers."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import idefics2 as I

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    GOLD = json.load(f)


def test_bucketed_position_ids_match_reference_source():
    for c in GOLD["idefics2_position_ids"]:
        m = np.asarray(c["patch_mask"]).astype(bool)
        got = I.bucketed_position_ids(m, c["num_patches_per_side"])
        assert got.tolist() == c["position_ids"]
    # the reference's `digitize(...) - 1` really produces negative buckets (coordinate 0 -> -1)
    assert min(min(r) for c in GOLD["idefics2_position_ids"] for r in c["position_ids"]) < 0


def test_patch_mask_and_padding_image_removal_match_reference_source():
    g = GOLD["idefics2_get_input_embeddings"]
    B, N, C, H, W = g["pixel_values_shape"]
    rng = np.random.default_rng(0)
    pv = rng.standard_normal((B, N, C, H, W)).astype(np.float32)
    pv[0, g["zero_image"]] = 0.0
    assert I.real_image_indices(pv) == [i for i in range(N) if i != g["zero_image"]]
    assert len(I.real_image_indices(pv)) == g["n_images_kept"]
    pam = np.zeros((N, H, W), bool)
    for i, (h, w) in enumerate(g["pixel_attention_valid"]):
        pam[i, :h, :w] = True
    keep = I.real_image_indices(pv)
    assert I.patch_attention_mask(pam[keep], 14).astype(int).tolist() == g["patch_mask"]
    assert np.allclose(g["pixel_sum"], g["pixel_sum_expected"], rtol=1e-4, atol=1e-3)


def test_idefics2_merge_matches_reference_source():
    cfg = I.tiny_cfg()
    cfg.image_token_index = 100
    for case in GOLD["idefics2_merge"]:
        ids = np.asarray(case["input_ids"])
        H, n = case["hidden"], case["n_feats"]
        feats = torch.from_numpy((1000 + np.arange(n * H, dtype=np.float32)).reshape(1, n, H))
        emb = torch.from_numpy(-(np.arange(ids.size * H, dtype=np.float32) + 1).reshape(1, ids.shape[1], H))
        if case["error"] is not None:
            with pytest.raises(ValueError, match="do not match"):
                I.merge(cfg, feats, emb, ids)
            continue
        assert np.array_equal(I.merge(cfg, feats, emb, ids).numpy(), np.asarray(case["output"], dtype=np.float32))


def _hf_model(c):
    pytest.importorskip("transformers")
    from transformers import Idefics2Config, Idefics2ForConditionalGeneration
    v, t, p = c.vision, c.text, c.perceiver
    cfg = Idefics2Config(
        vision_config=dict(hidden_size=v.hidden_size, num_hidden_layers=v.num_hidden_layers,
                           intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                           image_size=v.image_size, patch_size=v.patch_size, num_channels=v.num_channels,
                           layer_norm_eps=v.layer_norm_eps, hidden_act="quick_gelu"),
        perceiver_config=dict(hidden_act="silu", hidden_size=t.hidden_size, rms_norm_eps=t.rms_norm_eps,
                              resampler_n_latents=p.resampler_n_latents, resampler_depth=p.resampler_depth,
                              resampler_n_heads=p.resampler_n_heads, resampler_head_dim=p.resampler_head_dim,
                              num_key_value_heads=p.num_key_value_heads),
        text_config=dict(model_type="mistral", hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                         intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                         num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                         rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta, sliding_window=None,
                         max_position_embeddings=512, tie_word_embeddings=False),
        image_token_id=c.image_token_index, tie_word_embeddings=False)
    torch.manual_seed(0)
    return Idefics2ForConditionalGeneration(cfg).eval().float()


def _load_into_hf(m, W):
    sd = m.state_dict()
    new, used = {}, set()
    for k in sd:
        kk = k
        if k.startswith("model.vision_model."):
            kk = k[len("model."):]
        elif k.startswith("model.connector."):
            kk = k[len("model."):]
        elif k.startswith("model.text_model."):
            kk = "language_model." + k[len("model.text_model."):]
        elif k.startswith("lm_head."):
            kk = "language_model." + k
        if kk not in W:
            new[k] = sd[k]
            continue
        x = W[kk]
        if kk.endswith("patch_embedding.weight"):
            x = x.permute(0, 3, 1, 2).contiguous()
        assert tuple(x.shape) == tuple(sd[k].shape), (k, x.shape, sd[k].shape)
        new[k] = x.clone()
        used.add(kk)
    missing = sorted(set(W) - used)
    assert not missing, missing[:6]
    m.load_state_dict(new)


def test_idefics2_oracle_f32_matches_hf_transformers():
    c = I.tiny_cfg()
    try:
        m = _hf_model(c)
    except Exception as e:  # config API drift between transformers versions
        pytest.skip(f"cannot build the HF model here: {e}")
    W = I.init_weights(c, seed=5)
    _load_into_hf(m, W)
    req = I.synthetic_request(c, n_images=2, n_text=8, seed=2)
    ids, pv = req["input_ids"], req["pixel_values"]
    # HF's buckets (torch.bucketize, no "- 1") and eps for the wiring check; full images
    side = c.vision.image_size // c.vision.patch_size
    bound = np.linspace(1 / side, 1.0, side, endpoint=False)
    frac = np.linspace(0, 1, side, endpoint=False)
    bk = np.asarray(torch.bucketize(torch.from_numpy(frac), torch.from_numpy(bound), right=True))
    hf_pos = np.tile((bk[:, None] * side + bk).flatten()[None], (2, 1))
    out = I.greedy_generate(c, W, ids, pv, None, 1, dtype="f32", vision_dtype="f32",
                            position_ids=hf_pos, post_ln_eps=c.vision.layer_norm_eps)
    with torch.no_grad():
        hf = m(input_ids=torch.from_numpy(ids), pixel_values=torch.from_numpy(pv),
               attention_mask=torch.ones_like(torch.from_numpy(ids)))
    want, got = hf.logits[0, -1].float(), out["logits"][0][0]
    rel = float((got - want).norm() / want.norm())
    print(f"Idefics2 oracle f32 vs HF logits rel_l2={rel:.3e}")
    assert rel <= 5e-5
