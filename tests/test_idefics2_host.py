"""CPU: the host-side integer logic of the Idefics2 product path (mlx_vlm_b200/models/idefics2) against
the goldens generated from the reference's own source (tests/golden/make_golden.py): bucketed position
ids (incl. the negative buckets), pixel mask -> patch mask, padding-image removal.  Bit-exact."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    GOLD = json.load(f)


def test_bucketed_position_ids():
    from mlx_vlm_b200.models.idefics2.vision import bucketed_position_ids
    for c in GOLD["idefics2_position_ids"]:
        m = np.asarray(c["patch_mask"]).astype(bool)
        assert bucketed_position_ids(m, c["num_patches_per_side"]).tolist() == c["position_ids"]


def test_patch_mask_and_padding_images():
    from mlx_vlm_b200.models.idefics2.idefics2 import patch_attention_mask, real_image_indices
    g = GOLD["idefics2_get_input_embeddings"]
    B, N, C, H, W = g["pixel_values_shape"]
    rng = np.random.default_rng(0)
    pv = rng.standard_normal((B, N, C, H, W)).astype(np.float32)
    pv[0, g["zero_image"]] = 0.0
    keep = real_image_indices(pv)
    assert keep == [i for i in range(N) if i != g["zero_image"]] and len(keep) == g["n_images_kept"]
    pam = np.zeros((N, H, W), bool)
    for i, (h, w) in enumerate(g["pixel_attention_valid"]):
        pam[i, :h, :w] = True
    assert patch_attention_mask(pam[keep], 14).astype(int).tolist() == g["patch_mask"]


def test_config_defaults_and_sanitize():
    from mlx_vlm_b200.models.idefics2 import Model, ModelConfig
    from mlx_vlm_b200.models.idefics2.config import idefics2_8b_config
    cfg = idefics2_8b_config()
    assert cfg.perceiver_config.resampler_n_latents == 64 and cfg.text_config.vocab_size == 32003
    d = ModelConfig.from_dict({"text_config": {"hidden_size": 64}, "vision_config": {"hidden_size": 32},
                               "perceiver_config": {"resampler_depth": 1}, "image_token_id": 7, "junk": 1})
    assert d.image_token_index == 7 and d.text_config.hidden_size == 64
    m = Model.__new__(Model)
    out = m.sanitize({"model.text_model.layers.0.x": 1, "lm_head.weight": 2, "model.vision_model.a": 3,
                      "model.connector.b": 4})
    assert set(out) == {"language_model.layers.0.x", "language_model.lm_head.weight", "vision_model.a", "connector.b"}


def test_prepare_inputs_passes_processor_extras_through():
    """utils.py:2124-2134: keys the processor adds (Idefics2: `pixel_attention_mask`, LLaVA-Next: `image_sizes`)
    reach the model as kwargs; `images` is an alias of `pixel_values` (utils.py:2113-2115)."""
    from mlx_vlm_b200.utils import prepare_inputs

    def proc(text=None, images=None, **kw):
        return {"input_ids": [[1, 2, 3]], "attention_mask": [[1, 1, 1]], "images": np.zeros((1, 2, 3, 4, 4), np.float32),
                "pixel_attention_mask": np.ones((1, 2, 4, 4), bool), "image_sizes": [[4, 4]], "note": "x"}
    out = prepare_inputs(proc, images=[np.zeros((4, 4, 3), np.uint8)], prompts="hi", device="cpu")
    assert out["input_ids"].tolist() == [[1, 2, 3]] and out["pixel_values"].shape == (1, 2, 3, 4, 4)
    assert out["pixel_attention_mask"].dtype == bool and out["pixel_attention_mask"].shape == (1, 2, 4, 4)
    assert out["image_sizes"] == [[4, 4]] and out["note"] == "x" and "images" not in out
