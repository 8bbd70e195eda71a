"""GPU parity of the whole generate path (vision tower -> merge -> rope index ->
prefill -> decode -> greedy sampling) against the oracle on the same seeded
weights and inputs, through the reference-shaped Python surface
(Model.get_input_embeddings / language_model / generate_step) which calls the C ABI.

Bars: integer outputs (input ids, merge indices, position ids, rope deltas)
bit-exact; bf16 logits relative L2 <= 1e-3 (north star) — measured here with the
oracle's rounding points mirrored, typically ~1e-4; greedy token ids equal where
the oracle's top-2 logprob margin is non-zero.
"""
import numpy as np
import pytest
import torch

from _util import cmp_bf16

pytestmark = pytest.mark.gpu


def _mk_cfg(kind):
    from oracle import qwen2vl as O
    if kind == "tiny":
        return O.tiny_cfg()
    if kind == "wide2":  # real Qwen2-VL-2B widths, 2 LM layers + 2 ViT blocks
        c = O.qwen2_vl_2b()
        c.text.num_hidden_layers = 2
        c.vision.depth = 2
        return c
    if kind == "full":
        return O.qwen2_vl_2b()
    raise ValueError(kind)


def _to_model_config(c):
    from mlx_vlm_b200.models.qwen2_vl.config import ModelConfig, TextConfig, VisionConfig
    t, v = c.text, c.vision
    text = TextConfig(model_type="qwen2_vl", hidden_size=t.hidden_size,
                      num_hidden_layers=t.num_hidden_layers, intermediate_size=t.intermediate_size,
                      num_attention_heads=t.num_attention_heads, rms_norm_eps=t.rms_norm_eps,
                      vocab_size=t.vocab_size, num_key_value_heads=t.num_key_value_heads,
                      rope_theta=t.rope_theta,
                      rope_scaling={"type": "mrope", "mrope_section": list(t.mrope_section)},
                      tie_word_embeddings=t.tie_word_embeddings)
    vis = VisionConfig(depth=v.depth, embed_dim=v.embed_dim, hidden_size=v.hidden_size,
                       num_heads=v.num_heads, patch_size=v.patch_size, mlp_ratio=v.mlp_ratio,
                       spatial_merge_size=v.spatial_merge_size,
                       temporal_patch_size=v.temporal_patch_size)
    return ModelConfig(text_config=text, vision_config=vis, model_type="qwen2_vl",
                       image_token_id=c.image_token_id, video_token_id=c.video_token_id,
                       vision_start_token_id=c.vision_start_token_id, vocab_size=t.vocab_size)


def _build(kind, n_text, hw, seed=0, jitter=0.05):
    from mlx_vlm_b200.models.qwen2_vl import Model
    from oracle import qwen2vl as O
    c = _mk_cfg(kind)
    W = O.init_weights(c, seed, norm_jitter=jitter)
    model = Model(_to_model_config(c), device="cuda:0")
    model.load_weights(W)
    req = O.synthetic_request(c, n_text, image_hw=hw, seed=seed)
    return c, W, model, req


@pytest.mark.parametrize("kind,n_text,hw,n_dec", [("tiny", 12, (56, 84), 12),
                                                   ("wide2", 32, (112, 112), 8)])
def test_generate_path_parity(kind, n_text, hw, n_dec):
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.cache import make_prompt_cache
    from oracle import qwen2vl as O
    c, W, model, req = _build(kind, n_text, hw)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    eng = model.engine
    # ---------------- oracle: free-running greedy + teacher-forced replay
    ref = O.greedy_generate(c, W, ids, pv, grid, n_dec)
    pre = ref["prefill"]
    # ---------------- vision tower
    pvd = torch.from_numpy(pv).cuda()
    feats = model.vision_tower(pvd, grid)
    eng.stream.synchronize()
    cmp_bf16(feats, pre.image_features, f"{kind} vision features", max_mismatch=0.05)
    # ---------------- input embeddings: merge indexing, rope index (bit-exact)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    eng.stream.synchronize()
    assert np.array_equal(np.asarray(emb.position_ids), pre.position_ids)
    assert np.array_equal(np.asarray(emb.rope_deltas), pre.rope_deltas)
    src = O.merge_indices(c, ids)[0]
    e = emb.inputs_embeds[0].float().cpu()
    table = W["language_model.model.embed_tokens.weight"]
    text_rows = np.where(src < 0)[0]
    assert torch.equal(e[text_rows], table[torch.from_numpy(ids[0][text_rows])])
    img_rows = np.where(src >= 0)[0]
    assert torch.equal(e[img_rows], feats.float().cpu()[torch.from_numpy(src[img_rows])]), \
        "image rows must be exact copies of the feature rows selected by cumsum(mask)-1"
    # ---------------- prefill: all-row logits through LanguageModel.__call__
    cache = make_prompt_cache(model.language_model)
    out = model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=cache,
                               position_ids=emb.position_ids, rope_deltas=emb.rope_deltas)
    eng.stream.synchronize()
    T = ids.shape[1]
    assert out.logits.shape == (1, T, c.text.vocab_size) and cache[0].offset == T
    cmp_bf16(out.logits[0, -1], pre.logits_last[0], f"{kind} prefill logits (last row)",
             max_mismatch=0.2)
    cmp_bf16(eng.logits_view(), pre.logits_last[0], f"{kind} fused head logits", max_mismatch=0.2)
    cmp_bf16(eng.logprobs_view(), ref["logprobs"][0][0], f"{kind} logprobs", max_mismatch=0.2)
    # KV cache content (K rotated before caching, language.py:97-114)
    k_or = ref["cache"][0].keys[0, :, :T]
    cmp_bf16(cache[0].keys[0, :, :T], k_or, f"{kind} layer-0 K cache", max_mismatch=0.05)
    # ---------------- decode, teacher-forced with the oracle's tokens
    toks = ref["tokens"][0].tolist()
    delta = int(pre.rope_deltas[0, 0])
    eng.set_next(toks[0], T, T + delta)
    for n in range(1, n_dec):
        eng.decode(1, force_tokens=np.asarray([toks[n]], dtype=np.int32))
        eng.stream.synchronize()
        cmp_bf16(eng.logits_view(), ref["logits"][n][0], f"{kind} decode step {n} logits",
                 max_mismatch=0.2)
    # ---------------- free-running greedy through generate_step (public API)
    got, lps = [], []
    for tok, lp in generate_step(ids, model, pvd, None, max_tokens=n_dec, image_grid_thw=grid):
        got.append(tok)
        lps.append(lp)
    for n, (g, w) in enumerate(zip(got, toks)):
        lp = ref["logprobs"][n][0]
        top2 = torch.topk(lp, 2).values
        if float(top2[0] - top2[1]) > 0:
            assert g == w, f"token {n}: got {g}, oracle {w}"
        else:
            break  # a tie at the maximum: sequences may legitimately diverge after it
    cmp_bf16(lps[0], ref["logprobs"][0][0], f"{kind} yielded logprobs[0]", max_mismatch=0.2)


def test_text_only_and_cache_reuse():
    """text-only request (qwen2_vl.py:34-42) + chunked prefill == one-shot prefill
    (reference tests/cache_invariants.py:47-115)."""
    from mlx_vlm_b200.models.cache import make_prompt_cache
    from oracle import qwen2vl as O
    c, W, model, _ = _build("tiny", 8, (56, 56))
    eng = model.engine
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 900, size=(1, 37))
    emb = model.get_input_embeddings(ids, None)
    assert np.asarray(emb.position_ids).shape == (1, 37)
    assert int(np.asarray(emb.rope_deltas)[0, 0]) == 0
    ref = O.greedy_generate(c, W, ids, None, None, 1)
    c1 = make_prompt_cache(model.language_model)
    o1 = model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=c1,
                              position_ids=emb.position_ids)
    eng.stream.synchronize()
    cmp_bf16(o1.logits[0, -1], ref["prefill"].logits_last[0], "text-only logits", max_mismatch=0.2)
    # chunked: 20 + 17
    c2 = make_prompt_cache(model.language_model)
    model.language_model._position_ids = None
    model.language_model._rope_deltas = None
    model.language_model(ids[:, :20], inputs_embeds=emb.inputs_embeds[:, :20], cache=c2,
                         position_ids=emb.position_ids)
    o2 = model.language_model(ids[:, 20:], inputs_embeds=emb.inputs_embeds[:, 20:], cache=c2,
                              position_ids=emb.position_ids)
    eng.stream.synchronize()
    assert c2[0].offset == 37
    cmp_bf16(o2.logits[0, -1], ref["prefill"].logits_last[0], "chunked-prefill logits",
             max_mismatch=0.2)


def test_merge_count_mismatch_raises():
    from mlx_vlm_b200.models.qwen2_vl import Model
    c, W, model, req = _build("tiny", 8, (56, 56))
    feats = torch.zeros(3, c.text.hidden_size, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(ValueError, match="does not match"):
        Model.merge_input_ids_with_image_features(c.image_token_id, c.video_token_id, feats, None,
                                                  req["input_ids"], _engine=model.engine)


def test_full_size_c1():
    """BASELINE config C1: Qwen2-VL-2B shapes, 1 image 336x336, 32 text tokens (+144
    image tokens), 64 greedy tokens — oracle (CPU) vs CUDA path."""
    from mlx_vlm_b200.generate import generate_step
    from oracle import qwen2vl as O
    c, W, model, req = _build("full", 32, (336, 336), jitter=0.0)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    assert ids.shape[1] == 32 + 144
    n_dec = 64
    ref = O.greedy_generate(c, W, ids, pv, grid, n_dec, keep_logits=True)
    pvd = torch.from_numpy(pv).cuda()
    got, lps = [], []
    for tok, lp in generate_step(ids, model, pvd, None, max_tokens=n_dec, image_grid_thw=grid):
        got.append(tok)
        lps.append(lp)
    toks = ref["tokens"][0].tolist()
    n_cmp = 0
    for n in range(n_dec):
        if got[:n] != toks[:n]:
            break  # histories diverged (tie at a maximum earlier): stop comparing
        cmp_bf16(lps[n], ref["logprobs"][n][0], f"C1 logprobs step {n}", rel_l2=1e-3,
                 max_mismatch=0.3)
        n_cmp += 1
    print(f"C1: {n_cmp} steps compared; tokens equal: {got == toks}")
    assert n_cmp >= 8
    assert got[:n_cmp] == toks[:n_cmp]
