"""Independent cross-check of the oracle's WIRING (SURVEY §8c "HF-torch as tie-breaker"):
HuggingFace transformers' torch Qwen2VLForConditionalGeneration with the same random
weights, fp32, must give the same logits / vision features / M-RoPE position ids as the
oracle in its un-rounded ("f32") mode.  This pins structure (layer order, QKV split,
rotary pairing and M-RoPE sections, GQA, merge order, patch-embed flattening), not the
bf16 rounding points."""
import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _hf_model(c):
    from transformers import Qwen2VLConfig, Qwen2VLForConditionalGeneration
    t, v = c.text, c.vision
    cfg = Qwen2VLConfig(
        text_config=dict(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                         intermediate_size=t.intermediate_size,
                         num_attention_heads=t.num_attention_heads,
                         num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                         rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta,
                         tie_word_embeddings=True, max_position_embeddings=4096,
                         rope_scaling={"type": "mrope", "mrope_section": list(t.mrope_section)}),
        vision_config=dict(depth=v.depth, embed_dim=v.embed_dim, hidden_size=v.hidden_size,
                           num_heads=v.num_heads, mlp_ratio=int(v.mlp_ratio),
                           patch_size=v.patch_size, spatial_merge_size=v.spatial_merge_size,
                           temporal_patch_size=v.temporal_patch_size, in_channels=3),
        image_token_id=c.image_token_id, video_token_id=c.video_token_id,
        vision_start_token_id=c.vision_start_token_id, vision_end_token_id=c.vision_end_token_id,
        tie_word_embeddings=True)
    torch.manual_seed(0)
    return Qwen2VLForConditionalGeneration(cfg).eval().float()


def _weights_from_hf(m):
    W = {}
    for k, x in m.state_dict().items():
        if k.startswith("model.visual."):
            k = "vision_tower." + k[len("model.visual."):]
        elif k.startswith("model.language_model."):
            k = "language_model.model." + k[len("model.language_model."):]
        elif k.startswith("lm_head."):
            continue  # tied
        W[k] = x.detach().float().clone()
    return W


def test_oracle_f32_matches_hf_transformers():
    from oracle import qwen2vl as O
    c = O.tiny_cfg()
    try:
        m = _hf_model(c)
    except Exception as e:  # config API drift between transformers versions
        pytest.skip(f"cannot build the HF model here: {e}")
    W = _weights_from_hf(m)
    assert set(O.weight_shapes(c)) <= set(W), sorted(set(O.weight_shapes(c)) - set(W))[:5]
    for k, shp in O.weight_shapes(c).items():
        assert tuple(W[k].shape) == tuple(shp), k
    req = O.synthetic_request(c, 12, image_hw=(56, 84))
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    ref = O.greedy_generate(c, W, ids, pv, grid, 1, dtype="f32")
    tids = torch.from_numpy(ids)
    mm = (tids == c.image_token_id).long()
    with torch.no_grad():
        try:
            out = m(input_ids=tids, pixel_values=torch.from_numpy(pv),
                    image_grid_thw=torch.from_numpy(grid), mm_token_type_ids=mm)
        except TypeError:
            out = m(input_ids=tids, pixel_values=torch.from_numpy(pv),
                    image_grid_thw=torch.from_numpy(grid))
        feats = m.model.visual(torch.from_numpy(pv), grid_thw=torch.from_numpy(grid))
    feats = getattr(feats, "pooler_output", feats)
    if isinstance(feats, (tuple, list)):
        feats = feats[0]
    mine_feats = ref["prefill"].image_features
    assert feats.shape == mine_feats.shape
    ef = float((feats - mine_feats).norm() / feats.norm())
    hf, mine = out.logits[0, -1], ref["prefill"].logits_last[0]
    el = float((hf - mine).norm() / hf.norm())
    print(f"vision features rel err {ef:.2e}; logits rel err {el:.2e}")
    assert ef < 1e-4 and el < 1e-4
    assert int(hf.argmax()) == int(mine.argmax())
