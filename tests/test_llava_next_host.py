"""CPU: LLaVA-Next — the oracle (oracle/llava_next.py) and the product's host logic
(models/llava_next/llava_next.py::merge_plan) against the goldens produced by EXECUTING the reference's own
`get_input_embeddings` / `_merge_input_ids_with_image_features` (tests/golden/make_llava_next_golden.py):
crop selection, class-token drop, newline blocks along the crop axis, the growing merge and its zip() truncation."""
import json
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "llava_next_golden.json")) as f:
    GOLD = json.load(f)
V, TOK = 50, 32


def _stub_blocks(c):
    H, P, N = c["hidden"], c["patches"], c["n_crops"]
    layer = c["layer"] % 4
    nn = np.arange(N, dtype=np.float32)[:, None, None]
    t = np.arange(P + 1, dtype=np.float32)[None, :, None]
    ch = np.arange(H, dtype=np.float32)[None, None, :]
    st = 100.0 * (layer + 1) + 1000.0 * nn + 10.0 * t + ch        # the golden generator's tower
    sel = st[:, 1:] if c["strategy"] == "default" else st
    f = torch.from_numpy(2.0 * sel + 1.0)                           # ... and projector
    nl = torch.tensor(c["newline"])[None, None, :].expand_as(f)
    return torch.cat([f, nl], 0)


@pytest.mark.parametrize("case", [c for c in GOLD["cases"] if c["error"] is None], ids=lambda c: c["tag"][:40])
def test_merge_matches_reference_source(case):
    from oracle import llava_next as ON
    from mlx_vlm_b200.models.llava_next.llava_next import merge_plan
    ids = np.asarray(case["input_ids"])
    H = case["hidden"]
    table = torch.from_numpy(np.arange(V * H, dtype=np.float32).reshape(V, H) + 0.5)
    blocks = _stub_blocks(case)
    want = np.asarray(case["output"], dtype=np.float32)
    # oracle
    out = ON.merge(ON.LlavaNextCfg(image_token_index=TOK), blocks, table[torch.from_numpy(ids)], ids)
    assert np.array_equal(out.numpy(), want)
    # product host logic: the plan + "k-th image position takes the k-th feature row" (what b200_embed_merge does)
    plan, used = merge_plan(ids, TOK, blocks.shape[0], blocks.shape[1])
    flat = blocks[:max(used, 1)].reshape(-1, H).numpy()
    got, k = [], 0
    for t in plan:
        if t == TOK:
            got.append(flat[k])
            k += 1
        else:
            got.append(table[t].numpy())
    assert np.array_equal(np.stack(got)[None], want)


def test_unknown_strategy_and_config():
    from mlx_vlm_b200.models.llava_next import ModelConfig
    bad = [c for c in GOLD["cases"] if c["error"]]
    assert bad and "Unexpected feature selection strategy" in bad[0]["error"]
    cfg = ModelConfig.from_dict({"model_type": "llava_next", "text_config": {"model_type": "mistral"},
                                 "vision_config": {"model_type": "clip_vision_model"}, "image_token_index": 32000})
    assert cfg.text_config.num_key_value_heads == 8 and cfg.text_config.intermediate_size == 14336
    assert cfg.vision_config.image_size == 336 and cfg.vision_feature_layer == -2
