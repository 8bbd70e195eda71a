"""CPU: Idefics3 / SmolVLM — the oracle's and the product's integer logic against goldens produced by
EXECUTING the reference's own source (tests/golden/make_idefics3_golden.py): pixel shuffle, position ids (correct
buckets, ids written to the first n_valid positions, position embedding zeroed on padding patches); config rules."""
import json
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "idefics3_golden.json")) as f:
    GOLD = json.load(f)


def test_oracle_pixel_shuffle_matches_reference():
    from oracle import idefics3 as O3
    assert len(GOLD["pixel_shuffle"]) >= 3
    for c in GOLD["pixel_shuffle"]:
        seq, E = c["side"] ** 2, c["E"]
        x = torch.from_numpy((np.arange(seq, dtype=np.float32)[:, None] * 100 + np.arange(E, dtype=np.float32)[None, :])[None])
        out = O3.pixel_shuffle(x, c["scale"])
        assert list(out.shape) == c["out_shape"]
        assert np.array_equal(out[0].numpy(), np.asarray(c["out"], dtype=np.float32))


def test_position_embeddings_match_reference():
    from oracle import idefics3 as O3
    from mlx_vlm_b200.models.idefics3.vision import position_ids
    assert len(GOLD["embeddings"]) >= 4
    for e in GOLD["embeddings"]:
        gh, gw = e["grid"]
        table = np.asarray(e["table"], dtype=np.float32)
        want = np.asarray(e["out"], dtype=np.float32)
        if e["patch_mask"] is None:
            ids, m = O3.position_ids_and_mask(None, gh, gw, e["side"])
            assert ids is None and np.array_equal(table[np.arange(gh * gw)], want)
            continue
        pm = np.asarray(e["patch_mask"])[None].astype(bool)
        ids, m = O3.position_ids_and_mask(pm, gh, gw, e["side"])
        assert np.array_equal(table[ids[0]] * m[0][:, None], want)
        # the product: masked patches point at an appended all-zero row
        pid = position_ids(pm, e["side"], gh * gw)
        assert np.array_equal(pid, ids)
        pid[~pm.reshape(1, -1)] = table.shape[0]
        table0 = np.concatenate([table, np.zeros_like(table[:1])], 0)
        assert np.array_equal(table0[pid[0]], want)


def test_position_ids_where_float32_rounding_decides():
    """An image that fills the position grid makes every `k / n >= boundary` a tie: the ids then follow the float32
    arithmetic of `mx.arange` (oracle/idefics3.py::mlx_arange_f32).  Goldens: the reference's source over a stand-in whose
    float `arange` follows mlx's Metal kernel; oracle and product must agree with them and with each other."""
    from oracle import idefics3 as O3
    from mlx_vlm_b200.models.idefics3.vision import position_ids
    assert len(GOLD["position_ids"]) >= 5
    for c in GOLD["position_ids"]:
        side = c["side"]
        vh, vw = c["valid"]
        m = np.zeros((1, side, side), dtype=bool)
        m[0, :vh, :vw] = True
        ids, mask = O3.position_ids_and_mask(m, side, side, side)
        assert np.array_equal(ids[0] * mask[0], np.asarray(c["ids_times_mask"]))
        assert np.array_equal(position_ids(m, side, side * side), ids)
    # not the identity on a full 26 x 26 grid (HF's torch.bucketize gives the identity there): reproduced, not repaired
    full = np.ones((1, 26, 26), dtype=bool)
    assert not np.array_equal(O3.position_ids_and_mask(full, 26, 26, 26)[0][0], np.arange(676))


def test_configs_and_alias():
    from mlx_vlm_b200.models import idefics3, smolvlm
    from mlx_vlm_b200.models.idefics3.config import idefics3_8b_config
    c = idefics3_8b_config()
    assert c.image_token_index == 128257 and c.scale_factor == 2 and c.vision_config.image_size // c.vision_config.patch_size == 26
    d = idefics3.ModelConfig.from_dict({"text_config": {"hidden_size": 64, "num_key_value_heads": None, "num_attention_heads": 4},
                                        "vision_config": {"hidden_size": 32}, "image_token_id": 7, "junk": 1})
    assert d.image_token_index == 7 and d.text_config.num_key_value_heads == 4 and d.model_type == "idefics3"
    # SmolVLM's derived defaults (smolvlm/config.py:24-67)
    s = smolvlm.ModelConfig.from_dict({"text_config": {"hidden_size": 960, "head_dim": 64}, "vision_config": {"hidden_size": 768}})
    assert s.text_config.num_attention_heads == 15 and s.text_config.num_key_value_heads == 15
    assert (s.vision_config.num_attention_heads, s.vision_config.num_hidden_layers, s.vision_config.intermediate_size) == (12, 12, 3072)
    v = smolvlm.VisionConfig()
    assert (v.hidden_size, v.num_attention_heads, v.num_hidden_layers, v.intermediate_size) == (1152, 18, 27, 4304)
    assert smolvlm.VisionConfig(hidden_size=1000).num_attention_heads == 16 and smolvlm.TextConfig().num_attention_heads == 32
    assert issubclass(smolvlm.Model, idefics3.Model)
    m = idefics3.Model.__new__(idefics3.Model)
    out = m.sanitize({"model.text_model.layers.0.x": 1, "lm_head.weight": 2, "model.vision_model.a": 3, "model.connector.b": 4})
    assert set(out) == {"language_model.layers.0.x", "language_model.lm_head.weight", "vision_model.a", "connector.b"}


def test_oracle_gelu_tanh_against_torch():
    from oracle import mlx_semantics as S
    x = torch.linspace(-6, 6, 4001)
    got = S.gelu_tanh(S.Rounder("f32"), x)
    want = torch.nn.functional.gelu(x, approximate="tanh")
    assert torch.allclose(got, want, atol=2e-6)


def test_oracle_generates_with_and_without_images():
    from oracle import idefics3 as O3
    cfg = O3.Idefics3Cfg(vision=O3.I2.SiglipCfg(hidden_size=64, num_hidden_layers=1, intermediate_size=96, num_attention_heads=4,
                                                image_size=56, patch_size=14),
                         text=O3.I2.MistralCfg(hidden_size=128, num_hidden_layers=1, intermediate_size=256, num_attention_heads=2,
                                               num_key_value_heads=1, vocab_size=160), image_token_index=150)
    W = O3.init_weights(cfg, 0)
    pv = np.random.default_rng(0).standard_normal((1, 2, 3, 56, 56)).astype(np.float32)
    pv[0, 1] = 0
    ids = np.asarray([[5, 6] + [150] * 4 + [7, 8]])
    r = O3.greedy_generate(cfg, W, ids, pv, None, 2)
    assert r["image_features"].shape == (4, 128) and len(r["tokens"]) == 2
    assert len(O3.greedy_generate(cfg, W, np.asarray([[5, 6, 7]]), None, None, 2)["tokens"]) == 2
