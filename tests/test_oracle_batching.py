"""CPU: the lock-step batched formulation of the reference (left-padded rows, BatchKVCache write index,
per-row rope deltas; oracle/batching.py) gives the same result as every row alone — the statement the
round-1 time-multiplexed BatchGenerator relies on, and the acceptance oracle of a batched kernel."""
import numpy as np
import torch

from oracle import batching as OB
from oracle import qwen2vl as O


def _requests(c):
    rng = np.random.default_rng(11)
    r_img = O.synthetic_request(c, 9, image_hw=(56, 84), seed=3)
    r_img2 = O.synthetic_request(c, 5, image_hw=(56, 56), seed=4)
    return [dict(input_ids=r_img["input_ids"], pixel_values=r_img["pixel_values"], image_grid_thw=r_img["image_grid_thw"]),
            dict(input_ids=rng.integers(0, 900, size=(1, 7))),
            dict(input_ids=r_img2["input_ids"], pixel_values=r_img2["pixel_values"], image_grid_thw=r_img2["image_grid_thw"]),
            dict(input_ids=rng.integers(0, 900, size=(1, 21)))]


def test_left_padded_mask_matches_reference_golden():
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.json")))
    for c in g["causal_mask_left_padding"]:
        m = OB.left_padded_mask(c["N"], c["offset"], c["left_padding"]).int().numpy()
        assert m.tolist() == c["mask"]


def test_lock_step_batch_equals_rows_alone():
    c = O.tiny_cfg()
    W = O.init_weights(c, 2, norm_jitter=0.05)
    reqs = _requests(c)
    n_tok = 6
    for dtype, tol in (("f32", 2e-5), ("bf16", 3e-2)):
        out = OB.batched_greedy_generate(c, W, reqs, n_tok, dtype=dtype)
        assert out["left_padding"] == [max(len(r["input_ids"][0]) for r in reqs) - len(r["input_ids"][0]) for r in reqs]
        for b, r in enumerate(reqs):
            alone = O.greedy_generate(c, W, r["input_ids"], r.get("pixel_values"), r.get("image_grid_thw"), n_tok, dtype=dtype)
            assert int(out["rope_deltas"][b, 0]) == int(np.asarray(alone["prefill"].rope_deltas).reshape(-1)[0])
            for n in range(n_tok):
                a, w = out["logits"][n][b], alone["logits"][n][0]
                rel = float((a - w).norm() / w.norm())
                assert rel <= tol, (dtype, b, n, rel)
                if dtype == "f32":
                    assert int(out["tokens"][b, n]) == int(alone["tokens"][0, n]), (b, n)
                if int(out["tokens"][b, n]) != int(alone["tokens"][0, n]):
                    break  # bf16 near-tie: histories legitimately diverge
