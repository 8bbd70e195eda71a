"""CPU: Qwen2.5-VL — the oracle's and the product's integer logic (window permutation of the merge units,
window / frame boundaries, rotary ids, reverse permutation) against goldens produced by EXECUTING the reference's own
`VisionModel.get_window_index` / `rot_pos_emb` / `__call__` (tests/golden/make_qwen2_5_vl_golden.py); config rules."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "qwen2_5_vl_golden.json")) as f:
    GOLD = json.load(f)


def _check(get_window_index, segment_tables, rot_pos_ids, make_cfg):
    assert len(GOLD["cases"]) >= 7
    for c in GOLD["cases"]:
        order, raw = get_window_index(c["grid_thw"], c["window_size"], c["patch_size"], c["merge"])
        assert order.tolist() == c["window_index"] and raw.tolist() == c["cu_window_seqlens_raw"]
        order, windows, frames = segment_tables(c["grid_thw"], make_cfg(c))
        assert windows.tolist() == c["blocks"][0]["cu"] == c["blocks"][2]["cu"]       # windowed blocks
        assert frames.tolist() == c["blocks"][1]["cu"]                                # the full-attention block
        n, unit = int(frames[-1]), c["merge"] ** 2
        seen = np.arange(n).reshape(n // unit, unit)[order].reshape(-1)
        assert seen.tolist() == c["order"]
        assert seen.reshape(-1, unit)[np.argsort(order, kind="stable")].tolist() == c["out"]
        pos = rot_pos_ids(c["grid_thw"], c["merge"])
        assert np.array_equal(pos, np.asarray(c["rot_hw"]).reshape(n, 2, 2)[:, :, 0])
        pos_w = pos.reshape(n // unit, unit, 2)[order].reshape(n, 2)
        assert np.array_equal(pos_w, np.asarray(c["rot_in_window_order"]).reshape(n, 2, 2)[:, :, 0])


def test_oracle_window_logic_matches_reference():
    from oracle import qwen2_5vl as O
    _check(O.get_window_index, O.segment_tables, O.Q.rot_pos_ids,
           lambda c: O.VisionCfg(window_size=c["window_size"], patch_size=c["patch_size"], spatial_merge_size=c["merge"]))


def test_product_window_logic_matches_reference():
    from mlx_vlm_b200.models.qwen2_5_vl import VisionConfig
    from mlx_vlm_b200.models.qwen2_5_vl.vision import get_window_index, rot_pos_ids, segment_tables
    _check(get_window_index, segment_tables, rot_pos_ids,
           lambda c: VisionConfig(window_size=c["window_size"], patch_size=c["patch_size"], spatial_merge_size=c["merge"]))


def test_config_rules():
    from mlx_vlm_b200.models.qwen2_5_vl import Model, ModelConfig, TextConfig, VisionConfig
    from mlx_vlm_b200.models.qwen2_5_vl.config import qwen2_5_vl_3b_config
    v = VisionConfig()
    assert (v.hidden_size, v.intermediate_size, v.out_hidden_size, v.window_size, v.fullatt_block_indexes) == \
        (1280, 3420, 1536, 112, [7, 15, 23, 31])
    d = ModelConfig.from_dict({"model_type": "qwen2_5_vl", "hidden_size": 64, "num_hidden_layers": 2, "intermediate_size": 128,
                               "num_attention_heads": 4, "rms_norm_eps": 1e-6, "vocab_size": 99, "image_token_id": 7,
                               "rope_scaling": {"type": "mrope", "mrope_section": [2, 3, 3]}, "junk": 1,
                               "vision_config": {"depth": 2, "hidden_size": 32, "junk": 2}})
    assert d.text_config.hidden_size == 64 and d.text_config.num_key_value_heads == 4 and d.text_config.tie_word_embeddings
    assert d.text_config.mrope_section == [2, 3, 3] and d.vision_config.depth == 2 and d.image_token_id == 7
    with pytest.raises(ValueError):
        TextConfig(model_type="x", hidden_size=8, num_hidden_layers=1, intermediate_size=8, num_attention_heads=1,
                   rms_norm_eps=1e-6, vocab_size=8, rope_scaling={"type": "mrope"})
    with pytest.raises(ValueError):
        TextConfig(model_type="x", hidden_size=8, num_hidden_layers=1, intermediate_size=8, num_attention_heads=1,
                   rms_norm_eps=1e-6, vocab_size=8, rope_scaling={"type": "yarn", "mrope_section": [1]})
    c = qwen2_5_vl_3b_config()
    assert c.vision_config.out_hidden_size == c.text_config.hidden_size == 2048
    m = Model.__new__(Model)
    out = m.sanitize({"visual.blocks.0.norm1.weight": 1, "model.layers.0.x": 2, "lm_head.weight": 3})
    assert set(out) == {"vision_tower.blocks.0.norm1.weight", "language_model.model.layers.0.x", "language_model.lm_head.weight"}


def test_oracle_generates():
    from oracle import qwen2_5vl as O
    cfg = O.tiny_cfg()
    W = O.init_weights(cfg, 0)
    req = O.synthetic_request(cfg, 8, (8, 12))
    r = O.greedy_generate(cfg, W, req["input_ids"], req["pixel_values"], req["image_grid_thw"], 2)
    assert r["image_features"].shape == (24, 256) and len(r["tokens"]) == 2
    # windowed attention matters: with one window covering everything the features change
    cfg2 = O.tiny_cfg()
    cfg2.vision.window_size = 56 * 8
    r2 = O.greedy_generate(cfg2, W, req["input_ids"], req["pixel_values"], req["image_grid_thw"], 1)
    assert not np.allclose(r["image_features"].numpy(), r2["image_features"].numpy())
