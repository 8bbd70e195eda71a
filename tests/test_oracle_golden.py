"""CPU: pin the oracle (and the product's host-side integer logic) to the reference.

Sources of truth
  * tests/golden/reference_golden.json — outputs of the REFERENCE'S OWN FUNCTION SOURCE
    (extracted with `ast` from /root/reference and executed over a numpy stand-in for
    mlx.core; tests/golden/make_golden.py, provenance inside the JSON);
  * the known-answer vectors of the reference's own tests, transcribed below with
    file:line.
Integer outputs must be bit-exact; fp32 formulas within 1e-6.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mlx_semantics as S
from oracle import qwen2vl as O

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_golden.json")) as f:
    GOLD = json.load(f)


def _cfg100():
    return O.tiny_cfg(image_token_id=100, video_token_id=101, vision_start_token_id=99)


def _product_lm():
    from mlx_vlm_b200.models.qwen2_vl.config import ModelConfig, TextConfig, VisionConfig
    from mlx_vlm_b200.models.qwen2_vl.language import LanguageModel
    t = TextConfig(model_type="qwen2_vl", hidden_size=64, num_hidden_layers=1, intermediate_size=64,
                   num_attention_heads=2, rms_norm_eps=1e-6, vocab_size=128, num_key_value_heads=1)
    mc = ModelConfig(text_config=t, vision_config=VisionConfig(), model_type="qwen2_vl",
                     image_token_id=100, video_token_id=101, vision_start_token_id=99)
    return LanguageModel(t, mc, engine_getter=lambda: None)


# --------------------------------------------------------------------------
# reference tests/test_models.py:11866-11877 (TestMultiImageMRoPE), transcribed
_INPUT_IDS = [1, 2, 99, 100, 100, 100, 100, 5, 99, 100, 100, 100, 100, 7]
_GRID = [[1, 4, 4], [1, 4, 4]]
_EXPECTED_T = [0, 1, 2, 3, 3, 3, 3, 5, 6, 7, 7, 7, 7, 9]
_EXPECTED_H = [0, 1, 2, 3, 3, 4, 4, 5, 6, 7, 7, 8, 8, 9]
_EXPECTED_W = [0, 1, 2, 3, 4, 3, 4, 5, 6, 7, 8, 7, 8, 9]
_EXPECTED_DELTA = 9 + 1 - len(_INPUT_IDS)


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_reference_known_answer_two_image_prompt(impl):
    if impl == "oracle":
        fn = lambda ids, g, m=None: O.get_rope_index(_cfg100(), ids, g, None, m)
    else:
        lm = _product_lm()
        fn = lambda ids, g, m=None: lm.get_rope_index(np.asarray(ids), g, None, m)
    pos, d = fn([_INPUT_IDS], _GRID)
    assert pos.shape == (3, 1, len(_INPUT_IDS))
    assert pos[0, 0].tolist() == _EXPECTED_T
    assert pos[1, 0].tolist() == _EXPECTED_H
    assert pos[2, 0].tolist() == _EXPECTED_W
    assert int(np.asarray(d).reshape(-1)[0]) == _EXPECTED_DELTA
    # test_models.py:11904-11930: left padding keeps the valid suffix
    pad = 2
    pos, d = fn([[0] * pad + _INPUT_IDS], _GRID, [[0] * pad + [1] * len(_INPUT_IDS)])
    assert pos[0, 0, pad:].tolist() == _EXPECTED_T
    assert pos[1, 0, pad:].tolist() == _EXPECTED_H
    assert pos[2, 0, pad:].tolist() == _EXPECTED_W
    assert int(np.asarray(d).reshape(-1)[0]) == _EXPECTED_DELTA
    # tests/test_rope.py:30-60: a fully masked row yields delta 0
    pos, d = fn([[0, 0, 0, 0], [10, 99, 100, 11]], [[1, 2, 2]], [[0, 0, 0, 0], [1, 1, 1, 1]])
    assert pos.shape == (3, 2, 4) and np.asarray(d).shape == (2, 1)
    assert np.asarray(d).tolist()[0] == [0]


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_get_rope_index_matches_reference_source(impl):
    cfg = _cfg100()
    lm = _product_lm()
    for case in GOLD["get_rope_index"]["cases"]:
        args = (case["input_ids"], case["image_grid_thw"], case["video_grid_thw"], case["attention_mask"])
        if impl == "oracle":
            pos, d = O.get_rope_index(cfg, *args)
        else:
            pos, d = lm.get_rope_index(np.asarray(args[0]), *args[1:])
        assert np.array_equal(np.asarray(pos), np.asarray(case["position_ids"])), case["tag"]
        assert np.array_equal(np.asarray(d).reshape(-1), np.asarray(case["rope_deltas"]).reshape(-1)), case["tag"]


def test_merge_matches_reference_source():
    cfg = _cfg100()
    from mlx_vlm_b200.models.qwen2_vl import Model
    for case in GOLD["merge_input_ids_with_image_features"]:
        ids = np.asarray(case["input_ids"])
        B, T = ids.shape
        H, n = case["hidden"], case["n_feats"]
        feats = torch.from_numpy((1000 + np.arange(n * H, dtype=np.float32)).reshape(n, H))
        emb = torch.from_numpy(-(np.arange(B * T * H, dtype=np.float32) + 1).reshape(B, T, H))
        if case["error"] is not None:
            with pytest.raises(ValueError, match="does not match"):
                O.merge_input_ids_with_image_features(cfg, feats, emb, ids)
            # the product's host-side validation raises the same error before any GPU work
            with pytest.raises(ValueError, match="does not match"):
                Model.merge_input_ids_with_image_features(100, 101, feats, None, ids, _engine=None)
            continue
        out = O.merge_input_ids_with_image_features(cfg, feats, emb, ids)
        assert np.array_equal(out.numpy(), np.asarray(case["output"], dtype=np.float32)), case["tag"]
        # index form used by the CUDA kernel test
        src = O.merge_indices(cfg, ids)
        flat = np.where(src[..., None] >= 0, feats.numpy()[np.maximum(src, 0)], emb.numpy())
        assert np.array_equal(flat, np.asarray(case["output"], dtype=np.float32)), case["tag"]


def test_vision_rotary_matches_reference_source():
    v = O.VisionCfg()  # head_dim 80 -> rotary dim 40, as in the golden stub
    for case in GOLD["rot_pos_emb"]:
        fr = O.vision_rotary_freqs(case["grid_thw"], v).numpy()
        assert list(fr.shape) == case["freqs_shape"]
        assert abs(float(fr.astype(np.float64).sum()) - case["freqs_sum"]) <= 1e-3 * max(1.0, abs(case["freqs_sum"]))
        assert np.allclose(fr[:6], np.asarray(case["freqs_head"]), rtol=1e-6, atol=1e-6)
        assert np.allclose(fr[-3:], np.asarray(case["freqs_tail"]), rtol=1e-6, atol=1e-6)
    g = GOLD["apply_rotary_pos_emb_vision"]
    x, fr = torch.tensor(g["x"]), torch.tensor(g["freqs"])
    cos = torch.cos(fr).repeat(1, 2)[:, None, :][None]
    sin = torch.sin(fr).repeat(1, 2)[:, None, :][None]
    out = x * cos + O._rotate_half(x) * sin
    assert np.allclose(out.numpy(), np.asarray(g["out"]), rtol=1e-6, atol=1e-6)


def test_mrope_pieces_match_reference_source():
    g = GOLD["chunked_position_selector"]
    assert O.mrope_selector(g["mrope_section"], g["freq_dim"]).tolist() == g["selector"]
    assert O.mrope_selector([2, 3, 3], 8).tolist() == g["small"]
    g = GOLD["apply_rotary_embedding"]
    q, k, ang = torch.tensor(g["q"]), torch.tensor(g["k"]), torch.tensor(g["angles"])
    R = S.Rounder("f32")
    qe, ke = O.apply_mrope(R, q, k, torch.cos(ang)[:, 0], torch.sin(ang)[:, 0])
    assert np.allclose(qe.numpy(), np.asarray(g["q_out"]), rtol=1e-6, atol=1e-6)
    assert np.allclose(ke.numpy(), np.asarray(g["k_out"]), rtol=1e-6, atol=1e-6)
    # the product config exposes the same section default / override
    from mlx_vlm_b200.models.qwen2_vl.config import qwen2_vl_2b_config
    assert qwen2_vl_2b_config().text_config.mrope_section == [16, 24, 24]


def test_causal_mask_and_kvcache_match_reference_source():
    for c in GOLD["create_causal_mask"]:
        N, off = c["N"], c["offset"]
        Sk = off + N
        mine = (np.arange(Sk - N, Sk)[:, None] >= np.arange(Sk)[None]).astype(np.int32)
        assert mine.tolist() == c["mask"]  # bottom-right aligned, as oracle sdpa / the kernels use
    from mlx_vlm_b200.models.cache import KVCache
    rng = np.random.default_rng(0)
    for make in (O.OracleKVCache, KVCache):
        cache = make()
        trace = GOLD["KVCache_trace"]
        for step in trace:
            if "L" in step:
                kk = torch.randn(1, 2, step["L"], 4)
                ks, vs = cache.update_and_fetch(kk, kk + 1)
                assert int(cache.offset) == step["offset"]
                assert int(ks.shape[2]) == step["returned_len"]
                assert torch.equal(ks[..., -1, :], kk[..., -1, :])
                if make is O.OracleKVCache:
                    assert int(cache.keys.shape[2]) == step["capacity"]  # 256-step growth
            else:
                n = cache.trim(step["trim"])
                assert int(n) == step["trimmed"] and int(cache.offset) == step["offset"]


def test_preprocessing_oracle_equals_product_and_token_estimate():
    """reference tests/test_utils.py:1207-1256: estimated image tokens == grid.prod()//merge^2."""
    from mlx_vlm_b200.models.qwen2_vl.processing_qwen2_vl import Qwen2VLImageProcessor
    rng = np.random.default_rng(1)
    ip = Qwen2VLImageProcessor(image_mean=O.OPENAI_CLIP_MEAN, image_std=O.OPENAI_CLIP_STD)
    for hw in ((336, 336), (56, 84), (200, 310), (28, 1000)):
        img = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
        out = ip([img])
        pv, grid = O.preprocess_image(img.transpose(2, 0, 1), O.VisionCfg())
        assert np.array_equal(out["pixel_values"], pv) and out["image_grid_thw"].tolist() == [grid]
        assert ip.num_image_tokens(*hw) == int(np.prod(grid)) // 4
    assert O.smart_resize(336, 336) == (336, 336)


def test_greedy_sampler_semantics():
    """logits - logsumexp in the logits dtype (ar.py:368) collapses near-ties; argmax takes
    the lowest index (sample_utils.py:63-64)."""
    R = S.Rounder("bf16")
    logits = R.r(torch.tensor([[3.5, 3.515625, 1.0, 3.515625]]))
    lp = O.logprobs_from_logits(R, logits)
    assert int(S.argmax_lowest(lp)[0]) in (0, 1)
    x = torch.tensor([[1.0, 5.0, 5.0, 2.0]])
    assert int(S.argmax_lowest(x)[0]) == 1
    from mlx_vlm_b200.sample_utils import greedy_sampler
    assert int(greedy_sampler(x)[0]) == 1


def test_load_path_key_renames_and_conv_layout_match_reference_source():
    """`Model.sanitize` (qwen2_vl.py:179-190) and the patch-embed layout rule
    (vision.py:9-25, 292-310).  The reference normalises the conv weight to MLX [O,T,H,W,C]; the
    engine wants HF [O,C,T,H,W] (same linear map on (C,T,H,W)-ordered pixel rows): both layouts
    must come out as the HF tensor."""
    from mlx_vlm_b200.models.qwen2_vl import Model
    from mlx_vlm_b200.models.qwen2_vl.vision import VisionModel
    m = Model.__new__(Model)
    for k, want in GOLD["sanitize_keys"].items():
        assert list(m.sanitize({k: 0}).keys()) == [want], k
    g = GOLD["vision_sanitize"]
    hf = torch.arange(int(np.prod(g["hf_shape"])), dtype=torch.float32).reshape(g["hf_shape"])
    mlx_layout = torch.tensor(g["mlx_layout"], dtype=torch.float32)
    assert list(mlx_layout.shape) == g["mlx_layout_shape"] and g["idempotent"]
    assert torch.equal(mlx_layout, hf.permute(0, 2, 3, 4, 1))          # what the reference stores
    vm = VisionModel.__new__(VisionModel)
    key = "vision_tower.patch_embed.proj.weight"
    for src in (hf, mlx_layout):
        out = vm.sanitize({key: src, "vision_tower.blocks.0.attn.position_ids": 1,
                           "vision_tower.blocks.0.attn.qkv.weight": torch.zeros(2, 2)})
        assert sorted(out.keys()) == g["kept_keys"]
        assert torch.equal(out[key].contiguous(), hf)


def test_image_processor_matches_reference_source_bit_exactly():
    """`Qwen3VLImageProcessor._process_one` (processing_qwen3_vl.py:302-354) executed
    from the reference's source: smart_resize, PIL bicubic, rescale/normalise, temporal duplication,
    merge-group-major patch order.  Oracle and product must give the same bytes."""
    import hashlib
    from mlx_vlm_b200.models.qwen2_vl.processing_qwen2_vl import Qwen2VLImageProcessor
    ip = Qwen2VLImageProcessor(image_mean=O.OPENAI_CLIP_MEAN, image_std=O.OPENAI_CLIP_STD)
    for c in GOLD["image_processor"]:
        img = np.random.default_rng(c["seed"]).integers(0, 256, size=(3, c["hw"][0], c["hw"][1]), dtype=np.uint8)
        pv, grid = O.preprocess_image(img, O.VisionCfg())
        assert list(grid) == c["grid"] and list(pv.shape) == c["shape"], c["hw"]
        assert hashlib.sha256(np.ascontiguousarray(pv.astype(np.float32)).tobytes()).hexdigest() == c["sha256"], c["hw"]
        out = ip([img.transpose(1, 2, 0)])
        assert out["image_grid_thw"].tolist() == [c["grid"]]
        got = np.ascontiguousarray(np.asarray(out["pixel_values"], dtype=np.float32))
        assert hashlib.sha256(got.tobytes()).hexdigest() == c["sha256"], c["hw"]


def test_language_model_position_bookkeeping():
    """SURVEY §8 a7 (language.py:404-518): prefill uses get_rope_index, chunks slice the stored ids,
    decode steps use cache_offset + rope_delta on all three axes (== oracle.decode_position_ids),
    explicit position ids longer than the chunk are sliced at the cache offset."""
    lm = _product_lm()
    ids = np.asarray([_INPUT_IDS])
    grid = _GRID
    L = ids.shape[1]
    # one-shot prefill
    pos, d0 = lm.resolve_position_ids(ids, 0, image_grid_thw=grid)
    assert pos.shape == (3, 1, L) and pos[0, 0].tolist() == _EXPECTED_T and d0 == _EXPECTED_DELTA
    # decode steps: offset + delta, identical on the three axes, same as the oracle
    for off in (L, L + 1, L + 7):
        p, d = lm.resolve_position_ids(np.asarray([[5]]), off)
        want = O.decode_position_ids(off, np.asarray([[_EXPECTED_DELTA]]), 1)
        assert np.array_equal(np.asarray(p), np.asarray(want)) and d == _EXPECTED_DELTA
    # an explicit per-row delta (continuous batching passes rope_deltas per call)
    p, _ = lm.resolve_position_ids(np.asarray([[5]]), 40, rope_deltas_kw=np.asarray([[-3]]))
    assert np.asarray(p).reshape(3).tolist() == [37, 37, 37]
    # chunked prefill: second chunk slices the ids stored by the first call of the request
    lm2 = _product_lm()
    full, _ = lm2.resolve_position_ids(ids, 0, image_grid_thw=grid)
    lm2._rope_deltas = None  # state at the time the reference takes this branch (chunk before decode)
    p2, _ = lm2.resolve_position_ids(ids[:, 6:], 6)
    assert np.array_equal(np.asarray(p2), np.asarray(full)[..., 6:])
    # explicit position ids for the whole prompt, chunk of 4 at offset 3
    p3, _ = lm2.resolve_position_ids(ids[:, 3:7], 3, position_ids=np.asarray(full))
    assert np.array_equal(np.asarray(p3), np.asarray(full)[..., 3:7])
    # text-only 2-D positions are broadcast to three axes
    lm3 = _product_lm()
    p4, d4 = lm3.resolve_position_ids(np.asarray([[1, 2, 3, 4]]), 0)
    assert np.asarray(p4).shape == (3, 1, 4) and np.asarray(p4)[1, 0].tolist() == [0, 1, 2, 3] and d4 == 0
