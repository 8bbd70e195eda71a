"""Independent cross-check of the sibling oracles' WIRING (SURVEY §8c "HF-torch as tie-breaker"): HuggingFace transformers'
torch implementations of Qwen2.5-VL and Idefics3 with the same random weights, in fp32, must give the same vision
features and last-token logits as oracle/qwen2_5vl.py and oracle/idefics3.py in their un-rounded ("f32") mode.  This pins
structure (window permutation and its inverse, windowed vs full attention blocks, RMSNorm / SwiGLU-with-bias blocks, 2-D
rotary pairing, merger; Idefics3: bucketed position ids, tanh-GELU, pixel-shuffle order, connector, merge), not the bf16
rounding points."""
import numpy as np
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def test_qwen2_5_vl_oracle_f32_matches_hf_transformers():
    from oracle import qwen2_5vl as O
    try:
        from transformers import Qwen2_5_VLConfig, Qwen2_5_VLForConditionalGeneration
    except Exception as e:
        pytest.skip(f"no Qwen2.5-VL in this transformers: {e}")
    c = O.tiny_cfg()
    t, v = c.text, c.vision
    try:
        cfg = Qwen2_5_VLConfig(
            text_config=dict(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                             intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                             num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size, rms_norm_eps=t.rms_norm_eps,
                             rope_theta=t.rope_theta, tie_word_embeddings=True, max_position_embeddings=4096,
                             rope_scaling={"type": "mrope", "mrope_section": list(t.mrope_section)}),
            vision_config=dict(depth=v.depth, hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                               out_hidden_size=v.out_hidden_size, num_heads=v.num_heads, patch_size=v.patch_size,
                               spatial_merge_size=v.spatial_merge_size, temporal_patch_size=v.temporal_patch_size,
                               in_channels=3, window_size=v.window_size, fullatt_block_indexes=list(v.fullatt_block_indexes),
                               hidden_act="silu"),
            image_token_id=c.image_token_id, video_token_id=c.video_token_id, vision_start_token_id=c.vision_start_token_id,
            vision_end_token_id=c.vision_end_token_id, tie_word_embeddings=True)
        torch.manual_seed(0)
        m = Qwen2_5_VLForConditionalGeneration(cfg).eval().float()
    except Exception as e:  # config API drift between transformers versions
        pytest.skip(f"cannot build the HF model here: {e}")
    W = {}
    for k, x in m.state_dict().items():
        if k.startswith("model.visual."):
            k = "vision_tower." + k[len("model.visual."):]
        elif k.startswith("model.language_model."):
            k = "language_model.model." + k[len("model.language_model."):]
        elif k.startswith("lm_head."):
            continue
        W[k] = x.detach().float().clone()
    for k, shp in O.weight_shapes(c).items():
        assert k in W and tuple(W[k].shape) == tuple(shp), k
    # ragged windows on both edges + a second image
    rng = np.random.default_rng(0)
    grids = [(8, 12), (6, 10)]
    K = v.in_channels * v.temporal_patch_size * v.patch_size ** 2
    pv = np.concatenate([rng.standard_normal((h * w, K)).astype(np.float32) for h, w in grids], 0)
    grid = np.asarray([[1, h, w] for h, w in grids], dtype=np.int64)
    ids = [5, 6, 7]
    for h, w in grids:
        ids += [c.vision_start_token_id] + [c.image_token_id] * (h * w // 4) + [c.vision_end_token_id]
    ids = np.asarray([ids + [8, 9, 10, 11]], dtype=np.int64)
    ref = O.greedy_generate(c, W, ids, pv, grid, 1, dtype="f32")
    tids = torch.from_numpy(ids)
    # The M-RoPE index is the reference's (pinned by its own known-answer tests, tests/test_oracle_golden.py) and is handed
    # to HF: transformers 5.x computes a different temporal index for the second image of a prompt (48 instead of 12
    # here), which is a property of that HF version, not of the model wiring this test pins.
    pos, _ = O.Q.get_rope_index(O._qcfg(c), ids, grid, None, None)
    with torch.no_grad():
        kw = dict(input_ids=tids, pixel_values=torch.from_numpy(pv), image_grid_thw=torch.from_numpy(grid),
                  position_ids=torch.from_numpy(np.asarray(pos)).long())
        try:
            out = m(**kw, mm_token_type_ids=(tids == c.image_token_id).long())
        except TypeError:
            out = m(**kw)
        feats = m.model.visual(torch.from_numpy(pv), grid_thw=torch.from_numpy(grid))
    feats = getattr(feats, "pooler_output", feats)
    if isinstance(feats, (tuple, list)):
        feats = feats[0]
    assert feats.shape == ref["image_features"].shape
    ef = _rel(ref["image_features"], feats)
    el = _rel(ref["logits"][0][0], out.logits[0, -1])
    print(f"qwen2.5-vl: vision features rel err {ef:.2e}; logits rel err {el:.2e}")
    assert ef < 1e-4 and el < 1e-4
    assert int(out.logits[0, -1].argmax()) == ref["tokens"][0]


def test_idefics3_oracle_f32_matches_hf_transformers():
    from oracle import idefics3 as O3
    try:
        from transformers import Idefics3Config, Idefics3ForConditionalGeneration
    except Exception as e:
        pytest.skip(f"no Idefics3 in this transformers: {e}")
    I2 = O3.I2
    c = O3.Idefics3Cfg(
        # a 7 x 7 position grid under 6 x 6 patches: k / 6 never ties with j / 7, so the bucket ids do not depend on
        # float32 rounding (on a grid the image fills they do, and HF's then differ from the reference's: see
        # tests/test_idefics3_host.py::test_position_ids_where_float32_rounding_decides)
        vision=I2.SiglipCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=96, num_attention_heads=4, image_size=98,
                            patch_size=14),
        text=I2.MistralCfg(hidden_size=128, num_hidden_layers=2, intermediate_size=256, num_attention_heads=4,
                           num_key_value_heads=2, vocab_size=320, rope_theta=10000.0), image_token_index=300)
    v, t = c.vision, c.text
    try:
        cfg = Idefics3Config(
            vision_config=dict(hidden_size=v.hidden_size, num_hidden_layers=v.num_hidden_layers,
                               intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                               image_size=v.image_size, patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps,
                               hidden_act="gelu_pytorch_tanh"),
            text_config=dict(model_type="llama", hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                             intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                             num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size, rms_norm_eps=t.rms_norm_eps,
                             rope_theta=t.rope_theta, tie_word_embeddings=False, max_position_embeddings=2048,
                             pad_token_id=0),
            image_token_id=c.image_token_index, scale_factor=c.scale_factor, tie_word_embeddings=False, pad_token_id=0)
        torch.manual_seed(0)
        m = Idefics3ForConditionalGeneration(cfg).eval().float()
    except Exception as e:
        pytest.skip(f"cannot build the HF model here: {e}")
    W = {}
    for k, x in m.state_dict().items():
        x = x.detach().float().clone()
        if k.startswith("model.vision_model."):
            k = k[len("model."):]
            if k.endswith("patch_embedding.weight"):
                x = x.permute(0, 2, 3, 1).contiguous()        # [O, C, kH, kW] -> [O, kH, kW, C]
        elif k.startswith("model.connector."):
            k = k[len("model."):]
        elif k.startswith("model.text_model."):
            k = "language_model." + k[len("model.text_model."):]
        elif k.startswith("lm_head."):
            k = "language_model." + k
        W[k] = x
    for k, shp in O3.weight_shapes(c).items():
        assert k in W and tuple(W[k].shape) == tuple(shp), (k, tuple(W[k].shape) if k in W else None, shp)
    rng = np.random.default_rng(1)
    side = 84
    pv = rng.standard_normal((1, 3, 3, side, side)).astype(np.float32)
    pv[0, 1] = 0.0                                             # a padding image
    # all pixels valid: with a ragged pixel mask the REFERENCE deviates from HF by construction (it writes the bucketed ids
    # to the first n_valid sequence positions, zeroes the position embedding of padding patches and runs the encoder
    # without an attention mask — idefics3/vision.py:128-141,176-183; reproduced by the oracle and pinned by
    # tests/golden/idefics3_golden.json), so only the unmasked case can be compared with HF.
    pam = np.ones((1, 3, side, side), dtype=bool)
    per = (side // 14 // c.scale_factor) ** 2
    ids = np.asarray([[5, 6, 7] + [c.image_token_index] * (2 * per) + [8, 9, 10]], dtype=np.int64)
    ref = O3.greedy_generate(c, W, ids, pv, pam, 1, dtype="f32")
    with torch.no_grad():
        out = m(input_ids=torch.from_numpy(ids), pixel_values=torch.from_numpy(pv),
                pixel_attention_mask=torch.from_numpy(pam))
        feats = m.model.get_image_features(torch.from_numpy(pv), torch.from_numpy(pam))
    feats = getattr(feats, "pooler_output", feats)
    if isinstance(feats, (tuple, list)):
        feats = feats[0]
    feats = feats.reshape(-1, feats.shape[-1])
    ef = _rel(ref["image_features"], feats)
    el = _rel(ref["logits"][0][0], out.logits[0, -1])
    print(f"idefics3: image features rel err {ef:.2e}; logits rel err {el:.2e}")
    assert ef < 1e-4 and el < 1e-4
    assert int(out.logits[0, -1].argmax()) == ref["tokens"][0]
