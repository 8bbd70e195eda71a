"""Round-2 gate tests (VERDICT r1 "Next round" item 1): the configurations that were benchmarked or
claimed but had no parity test.

  * C2 exactly as benched: Qwen2-VL-2B shapes, T = 128 + 144 = 272 prompt rows, 511 teacher-forced
    decode steps (context up to 783): logits at steps {1, 128, 256, 511} against the oracle, K/V rows
    of layer 0 (identical inputs -> the 1e-3 bar) and of the last layer (deep -> noise bar);
  * Qwen2-VL-7B decoder GEOMETRY (28 q / 4 kv heads, hidden 3584, inter 18944; 2 layers) through the
    DEFAULT decode path: one launch per step (the persistent kernel must fit it), parity vs oracle;
  * the public API on the GPU: generate() / stream_generate() incl. EOS stop, PromptCacheState
    text-suffix reuse vs a cold run, logits processors (token history seeded with the prompt,
    ar.py:357-361) and a non-greedy sampler (stream ordering of the torch sampler ops).
"""
import types

import numpy as np
import pytest
import torch

from _util import cmp_bf16, cmp_noise, rl2
from test_engine_gpu import _build, _mk_cfg, _to_model_config, _token_ok

pytestmark = pytest.mark.gpu


def _oracle_one_shot(O, c, W, embeds, pos, dtype):
    """The whole teacher-forced sequence in ONE causal pass (row-independent ops: identical, rounding
    point for rounding point, to the oracle's own step-by-step decode, and ~100x faster on a CPU)."""
    R = O.Rounder(dtype)
    cache = [O.OracleKVCache() for _ in range(c.text.num_hidden_layers)]
    hidden = O.lm_layers_forward(c, W, embeds, pos, cache, R)
    return R, hidden, cache


def test_c2_as_benched_teacher_forced():
    from mlx_vlm_b200.models.cache import make_prompt_cache
    from oracle import qwen2vl as O
    n_text, n_out = 128, 512
    c, W, model, req = _build("full", n_text, (336, 336), jitter=0.0)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    T = ids.shape[1]
    assert T == 272
    eng = model.engine
    rng = np.random.default_rng(11)
    forced = rng.integers(0, 20000, size=n_out - 1).astype(np.int64)      # f_0 .. f_510
    Rb = O.Rounder("bf16")
    embeds, feats, pos, deltas = O.get_input_embeddings(c, W, ids, pv, grid, Rb)
    delta = int(np.asarray(deltas).reshape(-1)[0])
    table = W["language_model.model.embed_tokens.weight"]
    seq = torch.cat([embeds, table[torch.from_numpy(forced)][None]], dim=1)   # (1, 783, H)
    dpos = (np.arange(len(forced)) + T + delta)[None, None, :].repeat(3, axis=0)
    allpos = np.concatenate([np.asarray(pos), dpos], axis=2)
    _, hid_b, cache_b = _oracle_one_shot(O, c, W, seq, allpos, "bf16")
    R32, hid_f, cache_f = _oracle_one_shot(O, c, W, seq, allpos, "f32")
    S_ = T + len(forced)
    assert cache_b[0].offset == S_ == 783

    # CUDA: the LM is fed the ORACLE's merged embeddings -> identical inputs for layer 0
    cache = make_prompt_cache(model.language_model)
    emb_dev = embeds.to(device="cuda", dtype=torch.bfloat16)
    model.language_model(ids, inputs_embeds=emb_dev, cache=cache, position_ids=np.asarray(pos),
                         rope_deltas=np.asarray(deltas), logits_to_keep=1, reserve_tokens=T + n_out + 1)
    eng.stream.synchronize()
    cmp_noise(eng.logits_view(), O.lm_head(c, W, hid_b[:, T - 1], Rb)[0],
              O.lm_head(c, W, hid_f[:, T - 1], R32)[0], "C2 prefill logits (step 0)")
    eng.set_next(int(forced[0]), T, T + delta)
    done = 0
    launches0 = eng.launch_count
    for step in (1, 128, 256, 511):
        k = step - done
        force = np.zeros(k, dtype=np.int32)
        avail = forced[done + 1: done + 1 + k]
        force[:len(avail)] = avail
        eng.decode(k, force_tokens=force)
        eng.stream.synchronize()
        done = step
        row = T + step - 1
        cmp_noise(eng.logits_view(), O.lm_head(c, W, hid_b[:, row], Rb)[0],
                  O.lm_head(c, W, hid_f[:, row], R32)[0], f"C2 decode step {step} (ctx {row + 1}) logits")
    assert eng.device_error() == 0
    assert (eng.launch_count - launches0 - 4) == 511, "one launch per decode step (+1 state launch per call)"
    L = c.text.num_hidden_layers
    cmp_bf16(cache[0].keys[0, :, :S_], cache_b[0].keys[0, :, :S_], "C2 K rows layer 0, ctx 783",
             rel_l2=1e-3, max_mismatch=0.02)
    cmp_bf16(cache[0].values[0, :, :S_], cache_b[0].values[0, :, :S_], "C2 V rows layer 0, ctx 783",
             rel_l2=1e-3, max_mismatch=0.02)
    cmp_noise(cache[L - 1].keys[0, :, :S_], cache_b[L - 1].keys[0, :, :S_], cache_f[L - 1].keys[0, :, :S_],
              "C2 K rows last layer, ctx 783")
    cmp_noise(cache[L - 1].values[0, :, :S_], cache_b[L - 1].values[0, :, :S_],
              cache_f[L - 1].values[0, :, :S_], "C2 V rows last layer, ctx 783")


def _cfg_7b_geometry():
    from oracle import qwen2vl as O
    c = O.tiny_cfg()
    t = c.text
    t.hidden_size, t.num_attention_heads, t.num_key_value_heads = 3584, 28, 4
    t.intermediate_size, t.num_hidden_layers, t.vocab_size = 18944, 2, 32000
    t.mrope_section = (16, 24, 24)
    t.tie_word_embeddings = False
    c.vision.hidden_size = 3584
    c.vision.depth = 1
    return c


def test_7b_geometry_runs_on_the_persistent_kernel():
    from mlx_vlm_b200.models.cache import make_prompt_cache
    from mlx_vlm_b200.models.qwen2_vl import Model
    from oracle import qwen2vl as O
    c = _cfg_7b_geometry()
    W = O.init_weights(c, 0, norm_jitter=0.05)
    model = Model(_to_model_config(c), device="cuda:0")
    model.load_weights(W)
    req = O.synthetic_request(c, 10, image_hw=(56, 56), seed=0)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    eng = model.engine
    n_dec = 6
    ref = O.greedy_generate(c, W, ids, pv, grid, n_dec)
    toks = ref["tokens"][0].tolist()
    ex = O.greedy_generate(c, W, ids, pv, grid, n_dec, dtype="f32", force_tokens=toks[:])
    pvd = torch.from_numpy(pv).cuda()
    cache = make_prompt_cache(model.language_model)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=cache,
                         position_ids=emb.position_ids, rope_deltas=emb.rope_deltas)
    T = ids.shape[1]
    eng.set_next(toks[0], T, T + int(ref["prefill"].rope_deltas[0, 0]))
    l0 = eng.launch_count
    for n in range(1, n_dec):
        eng.decode(1, force_tokens=np.asarray([toks[n]], dtype=np.int32))
        eng.stream.synchronize()
        cmp_noise(eng.logits_view(), ref["logits"][n][0], ex["logits"][n][0], f"7B-geometry decode step {n}")
    assert eng.device_error() == 0
    per_step = (eng.launch_count - l0) / (n_dec - 1)
    assert per_step == 2, f"default decode path must be ONE kernel per step (+1 state launch per call): {per_step}"


def _tiny_model_and_processor(n_text=12):
    c, W, model, req = _build("tiny", n_text, (56, 56))
    from mlx_vlm_b200.models.qwen2_vl.processing_qwen2_vl import SyntheticProcessor
    from mlx_vlm_b200.utils import StoppingCriteria
    proc = SyntheticProcessor(model.config, n_text_tokens=n_text, seed=0)
    proc.tokenizer.stopping_criteria = StoppingCriteria([], proc.tokenizer)
    return c, W, model, req, proc


def test_generate_and_stream_generate_on_gpu():
    from mlx_vlm_b200 import PromptCacheState, generate, stream_generate
    from mlx_vlm_b200.generate import generate_step
    c, W, model, req, proc = _tiny_model_and_processor()
    model.config.eos_token_id = []
    rng = np.random.default_rng(0)
    image = rng.integers(0, 256, size=(56, 56, 3), dtype=np.uint8)
    # --- generate() == generate_step on the same prepared inputs
    from mlx_vlm_b200.utils import prepare_inputs
    inp = prepare_inputs(proc, images=[image], prompts="describe", device=model.engine.device,
                         stream=model.engine.stream)
    want = [t for t, _ in generate_step(inp["input_ids"], model, inp["pixel_values"], None, max_tokens=10,
                                        image_grid_thw=inp["image_grid_thw"])]
    res = generate(model, proc, "describe", image=[image], max_tokens=10)
    assert res.generation_tokens == 10 and res.finish_reason == "length"
    assert res.prompt_tokens == inp["input_ids"].shape[1]
    streamed = [r.token for r in stream_generate(model, proc, "describe", image=[image], max_tokens=10)]
    assert streamed[:10] == want, (streamed, want)
    # --- EOS stop: make the 4th generated token an EOS
    res2 = generate(model, proc, "describe", image=[image], max_tokens=10, eos_tokens=[want[3]])
    first = want.index(want[3])
    assert res2.finish_reason == "stop" and res2.generation_tokens == first + 1, (res2, want)
    proc.tokenizer.stopping_criteria.reset([])
    # --- PromptCacheState: text-only turn 1, then turn 2 = turn-1 prompt + its answer + new text
    state = PromptCacheState()
    p1 = rng.integers(0, 900, size=(1, 21))
    t1 = [r.token for r in stream_generate(model, proc, "", input_ids=p1, max_tokens=6,
                                           prompt_cache_state=state)]
    t1 = t1[:6]
    assert state.token_ids == p1[0].tolist() + t1 and state.cache[0].offset >= 21 + 5
    p2 = np.asarray([p1[0].tolist() + t1 + rng.integers(0, 900, size=9).tolist()])
    warm = list(stream_generate(model, proc, "", input_ids=p2, max_tokens=7, prompt_cache_state=state))
    cold = list(stream_generate(model, proc, "", input_ids=p2, max_tokens=7))
    assert warm[-1].cached_tokens >= 21 + 5 and cold[-1].cached_tokens == 0
    assert [r.token for r in warm][:7] == [r.token for r in cold][:7]
    assert model.engine.device_error() == 0


def test_logits_processors_and_sampler_path_on_gpu():
    """Non-fused path of generate_step: repetition penalty with the reference's token history (the
    prompt is part of it) against a CPU evaluation on the oracle's logits, and a temperature > 0
    sampler that is deterministic by construction (top_k = 1) against greedy."""
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.sample_utils import make_repetition_penalty
    from oracle import qwen2vl as O
    c, W, model, req = _build("tiny", 12, (56, 56))
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    pvd = torch.from_numpy(pv).cuda()
    n = 8
    greedy = [t for t, _ in generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid)]
    samp = [t for t, _ in generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid,
                                        temperature=0.7, top_k=1, seed=3)]
    assert samp == greedy, (samp, greedy)
    got, got_lp = [], []
    for t, lp_dev in generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid,
                                   repetition_penalty=1.6, repetition_context_size=20):
        got.append(t)
        got_lp.append(float(lp_dev.float()[t]))
    # CPU evaluation: oracle logits teacher-forced with the CUDA tokens; history = prompt + fed tokens
    R = O.Rounder("bf16")
    embeds, feats, pos, deltas = O.get_input_embeddings(c, W, ids, pv, grid, R)
    cache = [O.OracleKVCache() for _ in range(c.text.num_hidden_layers)]
    hidden = O.lm_layers_forward(c, W, embeds, pos, cache, R)
    logits = O.lm_head(c, W, hidden[:, -1, :], R)
    proc = make_repetition_penalty(1.6, 20)
    hist = [int(t) for t in ids.reshape(-1)]
    effect = 0.0
    for i, tok in enumerate(got):
        lg = proc(hist, logits.clone())
        lp = O.logprobs_from_logits(R, lg)
        assert _token_ok(tok, lp[0]), f"penalised token {i}: got {tok}, oracle argmax {int(lp[0].argmax())}"
        want_lp = float(lp[0][tok])
        assert abs(got_lp[i] - want_lp) <= 0.03 + 0.03 * abs(want_lp), (i, got_lp[i], want_lp)
        effect = max(effect, abs(want_lp - float(O.logprobs_from_logits(R, logits)[0][tok])))
        hist.append(tok)
        e = W["language_model.model.embed_tokens.weight"][torch.tensor([tok])][:, None, :]
        p = O.decode_position_ids(cache[0].offset, deltas, 1)
        hidden = O.lm_layers_forward(c, W, e, p, cache, R)
        logits = O.lm_head(c, W, hidden[:, -1, :], R)
    assert effect > 0.1, "the penalty (history = prompt + fed tokens) must have changed the logprobs"


def test_fused_greedy_decode_hook_has_the_reference_contract():
    """Drive the model exactly as `GenerationBatch._fused_greedy_step` does (ar.py:1015-1042):
    `sampled = lm.fused_greedy_decode(inputs[:, None], cache=prompt_cache, **fwd_kwargs)` with
    fwd_kwargs = {"rope_deltas": (B, 1)}; (B,) token ids come back; feeding them straight back in
    is the steady state.  Tokens must equal the public generate_step's."""
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.cache import make_prompt_cache
    c, W, model, req = _build("tiny", 12, (56, 56))
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    pvd = torch.from_numpy(pv).cuda()
    lm, eng = model.language_model, model.engine
    want = [t for t, _ in generate_step(ids, model, pvd, None, max_tokens=7, image_grid_thw=grid)]
    cache = make_prompt_cache(lm)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    out = lm(ids, inputs_embeds=emb.inputs_embeds, cache=cache, position_ids=emb.position_ids,
             rope_deltas=emb.rope_deltas)
    eng.stream.synchronize()
    # the token the fused head + sampler of the prefill call picked (bf16 logprobs, lowest index wins)
    first = int(eng.token_log_view()[(eng.tokens_launched - 1) % eng.token_log_capacity])
    assert first == want[0]
    fwd_kwargs = {"rope_deltas": np.asarray(emb.rope_deltas)}
    inputs = np.asarray([first])
    got = [first]
    for _ in range(6):
        sampled = lm.fused_greedy_decode(inputs[:, None], cache=cache, **fwd_kwargs)
        assert sampled is not None and tuple(sampled.shape) == (1,)
        eng.stream.synchronize()
        got.append(int(sampled[0]))
        inputs = sampled          # the reference feeds `_next_tokens` straight back
    assert got == want, (got, want)
    assert cache[0].offset == ids.shape[1] + 6


def test_samplers_on_the_device():
    """SURVEY §8 a13 on the GPU: (i) the sampler masks (top-k / top-p / min-p / top-n-sigma / p-less / typical-p) on
    DEVICE tensors against the goldens produced by the reference's own functions (tests/golden); (ii) the sampled
    generate path (temperature + top-p + top-k, seeded): every sampled token lies inside the oracle's nucleus for
    its step (teacher-forced on the tokens actually sampled), the yielded logprobs are the engine's, the run is
    reproducible under the same seed and the torch ops run on the engine's stream (no cross-stream race)."""
    import json
    import os
    from mlx_vlm_b200 import sample_utils as SU
    from mlx_vlm_b200.generate import generate_step
    from oracle import qwen2vl as O
    from oracle.mlx_semantics import Rounder
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_golden.json")) as f:
        g = json.load(f)["sampler_masks"]

    def arr(x):
        return torch.tensor(np.array([[float(v) for v in row] for row in x], dtype=np.float32), device="cuda")

    def same(a, want):
        w = arr(want)
        assert a.is_cuda and torch.equal(torch.isinf(a), torch.isinf(w))
        fin = ~torch.isinf(w)
        assert torch.allclose(a[fin], w[fin], atol=1e-6)
    lp, lg = arr(g["logprobs"]), arr(g["logits"])
    same(SU.apply_top_k(lp, 3), g["top_k_3"])
    same(SU.apply_top_p(lp, 0.7), g["top_p_0.7"])
    same(SU.apply_min_p(lp, 0.2), g["min_p_0.2"])
    same(SU.apply_min_p(lp, 0.6, 3), g["min_p_0.6_keep3"])
    same(SU.apply_top_n_sigma(lg, 1.0), g["top_n_sigma_1.0"])
    same(SU.apply_p_less(lg, 0.8), g["p_less_t0.8"])
    same(SU.apply_typical_p(lp, 0.6), g["typical_p_0.6"])

    c, W, model, req = _build("tiny", 10, (56, 56))
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    pvd = torch.from_numpy(pv).cuda()
    n, top_k, top_p, temp = 6, 40, 0.9, 1.3

    def run(seed):
        out = []
        for tok, lp_ in generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid, temperature=temp,
                                      top_p=top_p, top_k=top_k, seed=seed):
            out.append((int(tok), lp_.float().cpu()))
        return out
    a, b, other = run(11), run(11), run(12)
    assert [t for t, _ in a] == [t for t, _ in b], "same seed, same tokens"
    assert [t for t, _ in a] != [t for t, _ in other] or True   # (a different seed may coincide on a tiny vocabulary)
    toks = [t for t, _ in a]
    ref = O.greedy_generate(c, W, ids, pv, grid, n, force_tokens=toks)
    for i, (tok, lp_got) in enumerate(a):
        lp_ref = ref["logprobs"][i][0].float()
        assert rl2(lp_got, lp_ref) < 2e-2, f"step {i}: yielded logprobs are the model's logprobs"
        # the oracle's nucleus for this step (same filters on the oracle's logprobs), with a tolerance band
        keep = SU.apply_top_k(SU.apply_top_p(lp_ref[None], top_p), top_k)[0]
        floor = keep[torch.isfinite(keep)].min()
        tol = 0.1   # the engine's logprobs sit within bf16 noise of the oracle's: a boundary token may fall either way
        assert float(lp_ref[tok]) >= float(floor) - tol, f"step {i}: sampled token {tok} is outside the oracle's nucleus"
    assert model.engine.device_error() == 0
