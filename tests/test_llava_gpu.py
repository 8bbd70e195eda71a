"""LLaVA-1.5 product path on the GPU against oracle/llava.py — the reference's
semantics: CLIP tower + projector in fp32 with bf16-valued weights, ONE rounding at the merge, Llama LM in
bf16.  The tower ops are checked on their own against fp32 torch first (identical inputs: 1e-5), then the
features at tiny and at CLIP-L widths, the merge (bit-exact copies) and the generate path."""
import numpy as np
import pytest
import torch

from _util import cmp_noise, rl2
from test_engine_gpu import _token_ok

pytestmark = pytest.mark.gpu


def _model_cfg(c):
    from mlx_vlm_b200.models.llava.config import ModelConfig, TextConfig, VisionConfig
    v, t = c.vision, c.text
    return ModelConfig(
        text_config=TextConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                               intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                               num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                               rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta),
        vision_config=VisionConfig(num_hidden_layers=v.num_hidden_layers, hidden_size=v.hidden_size,
                                   intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                                   image_size=v.image_size, patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps),
        image_token_index=c.image_token_index, vision_feature_layer=c.vision_feature_layer,
        vision_feature_select_strategy=c.vision_feature_select_strategy, vocab_size=t.vocab_size)


def _build(c, seed=0):
    from mlx_vlm_b200.models.llava import Model
    from oracle import llava as OL
    W = OL.init_weights(c, seed)
    model = Model(_model_cfg(c), device="cuda:0")
    model.load_weights(W)
    return W, model


def test_tower_ops_against_fp32_torch():
    from mlx_vlm_b200.models.tower_ops import EPI_GELU_EXACT, EPI_GELU_FAST, SplitBuf, TowerOps
    from mlx_vlm_b200.models.llava import Model
    from oracle import llava as OL
    model = Model(_model_cfg(OL.tiny_cfg()), device="cuda:0")
    eng = model.engine
    ops = TowerOps(eng)
    g = torch.Generator(device="cuda").manual_seed(0)
    with torch.cuda.stream(eng.stream):
        T, K, Nn = 577, 1024, 1536
        x = torch.randn(T, K, device="cuda", generator=g) * 3
        w = (torch.randn(Nn, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        b = (torch.randn(Nn, device="cuda", generator=g)).to(torch.bfloat16)
        res = torch.randn(T, Nn, device="cuda", generator=g)
        xs = SplitBuf(eng, T, K)
        ops.split(x, xs)
        out = ops.f32(T, Nn)
        ops.linear(xs, w, b, out32=out, res32=res, epi=EPI_GELU_FAST)
        lin = x.double() @ w.double().t() + b.double()
        want = (lin * torch.sigmoid(1.702 * lin) + res.double()).float()
        eng.stream.synchronize()
        e = rl2(out, want)
        print(f"split GEMM + bias + quick-gelu + residual: rel={e:.2e}")
        assert e < 2e-5
        mid = SplitBuf(eng, T, Nn)
        ops.linear(xs, w, b, out_split=mid, epi=EPI_GELU_EXACT)
        eng.stream.synchronize()
        want2 = (lin * (1 + torch.erf(lin / 2 ** 0.5)) / 2).float()
        got2 = mid.t[:, :Nn].float() + mid.t[:, mid.n_pad:mid.n_pad + Nn].float()
        assert rl2(got2, want2) < 2e-5
        # K that is not a multiple of 64 (CLIP patch embedding: 588 columns, weight padded to 592)
        K2 = 588
        x2 = torch.randn(300, K2, device="cuda", generator=g)
        w2 = torch.zeros(256, 592, device="cuda", dtype=torch.bfloat16)
        w2[:, :K2] = (torch.randn(256, K2, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
        xs2 = SplitBuf(eng, 300, K2)
        ops.split(x2, xs2)
        out2 = ops.f32(300, 256)
        ops.linear(xs2, w2, None, out32=out2, k_w=K2)
        eng.stream.synchronize()
        assert rl2(out2, (x2.double() @ w2[:, :K2].double().t()).float()) < 2e-5
        # LayerNorm fp32
        lw = (1 + 0.1 * torch.randn(K, device="cuda", generator=g)).to(torch.bfloat16)
        lb = (0.1 * torch.randn(K, device="cuda", generator=g)).to(torch.bfloat16)
        o32 = ops.f32(T, K)
        osp = SplitBuf(eng, T, K)
        ops.layer_norm(x, lw, lb, 1e-5, out32=o32, out_split=osp)
        eng.stream.synchronize()
        wl = torch.nn.functional.layer_norm(x.double(), (K,), lw.double(), lb.double(), 1e-5).float()
        assert rl2(o32, wl) < 1e-5
        assert rl2(osp.t[:, :K].float() + osp.t[:, osp.n_pad:osp.n_pad + K].float(), wl) < 2e-5
        # attention fp32, 2 segments, CLIP head geometry, and a GQA / rectangular case (perceiver-like)
        for (nh, nkv, hd, Lq, S, nseg) in ((16, 16, 64, 577, 577, 2), (16, 4, 96, 64, 200, 1), (4, 4, 16, 10, 10, 3),
                                           (16, 16, 72, 130, 130, 1)):
            E, Ek = nh * hd, nkv * hd
            q = torch.randn(nseg * Lq, E, device="cuda", generator=g)
            k = torch.randn(nseg * S, Ek, device="cuda", generator=g)
            v = torch.randn(nseg * S, Ek, device="cuda", generator=g)
            o = ops.f32(nseg * Lq, E)
            ops.attention((q, E, hd), (k, Ek, hd), (v, Ek, hd), n_heads=nh, n_kv=nkv, hd=hd, Lq=Lq, S=S, n_seg=nseg,
                          q_seg=Lq, k_seg=S, scale=hd ** -0.5, out32=o)
            eng.stream.synchronize()
            qd = q.double().view(nseg, Lq, nh, hd).transpose(1, 2)
            kd = k.double().view(nseg, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, dim=1)
            vd = v.double().view(nseg, S, nkv, hd).transpose(1, 2).repeat_interleave(nh // nkv, dim=1)
            wa = torch.softmax(qd @ kd.transpose(-1, -2) * hd ** -0.5, -1) @ vd
            wa = wa.transpose(1, 2).reshape(nseg * Lq, E).float()
            e = rl2(o, wa)
            print(f"attention_f32 nh={nh} nkv={nkv} hd={hd} Lq={Lq} S={S}: rel={e:.2e}")
            assert e < 2e-5


@pytest.mark.parametrize("kind", ["tiny", "clip_l_2layers", "clip_l_full_depth"])
def test_llava_features_merge_and_generate(kind):
    from oracle import llava as OL
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200.generate import generate_step
    # the decoder engine supports head_dim 64 / 128: text side 256 / 4 heads = 64
    c = OL.LlavaCfg(vision=OL.ClipCfg(hidden_size=64, num_hidden_layers=3, intermediate_size=128,
                                      num_attention_heads=4, image_size=42, patch_size=14),
                    text=OL.LlamaCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512,
                                     num_attention_heads=4, num_key_value_heads=2, vocab_size=320),
                    image_token_index=300)
    if kind == "clip_l_full_depth":   # the real CLIP-L/14-336 tower: 24 layers, 577 tokens (feature layer -2 = 23 run)
        c = OL.LlavaCfg(vision=OL.ClipCfg(),
                        text=OL.LlamaCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024,
                                         num_attention_heads=4, num_key_value_heads=4, vocab_size=33000),
                        image_token_index=32000)
    if kind == "clip_l_2layers":   # CLIP-L/14-336 widths, 3 encoder layers (feature layer -2 = after layer 1)
        c = OL.LlavaCfg(vision=OL.ClipCfg(num_hidden_layers=3),
                        text=OL.LlamaCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024,
                                         num_attention_heads=4, num_key_value_heads=4, vocab_size=33000),
                        image_token_index=32000)
    W, model = _build(c)
    eng = model.engine
    req = OL.synthetic_request(c, n_text=8, seed=1)
    ids, pv = req["input_ids"], req["pixel_values"]            # pv: NHWC fp32
    n = 4
    ref = OL.greedy_generate(c, W, ids, pv, n)
    ex = OL.greedy_generate(c, W, ids, pv, n, dtype="f32")
    pv_nchw = pv.permute(0, 3, 1, 2).contiguous().cuda()
    feats = model.encode_image(pv_nchw)
    eng.stream.synchronize()
    want_f32 = OL.image_features(c, W, pv, Rounder("f32"))      # the reference's fp32 tower, before the merge cast
    e = rl2(feats.float().cpu(), want_f32)
    mism = float((feats.float().cpu().reshape(-1) != ref["image_features"].reshape(-1)).float().mean())
    print(f"{kind}: features vs the fp32 tower rel={e:.2e} (bf16 rounding alone ~2e-3); "
          f"bf16 elements that differ from the oracle's rounding: {mism:.4f}")
    assert rl2(feats.float().cpu(), ref["image_features"]) < 1e-3
    assert mism < 0.02
    emb = model.get_input_embeddings(ids, pv_nchw)
    eng.stream.synchronize()
    pos = OL.merge_positions(c, ids)
    e_cpu = emb.inputs_embeds[0].float().cpu()
    assert torch.equal(e_cpu[pos], feats.float().cpu().reshape(-1, feats.shape[-1])), "image rows are pure copies"
    text_rows = [i for i in range(ids.shape[1]) if i not in set(pos)]
    table = W["language_model.model.embed_tokens.weight"]
    assert torch.equal(e_cpu[text_rows], table[torch.from_numpy(ids[0][text_rows])])
    with pytest.raises(ValueError):
        model._merge_input_ids_with_image_features(torch.zeros(1, len(pos) + 1, feats.shape[-1], dtype=torch.bfloat16,
                                                               device="cuda"), None, ids)
    toks = ref["tokens"]
    for i, (tok, lp) in enumerate(generate_step(ids, model, pv_nchw, None, max_tokens=n)):
        lp_ref = OL.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        assert _token_ok(tok, lp_ref), f"{kind}: token {i}: {tok} vs {toks[i]}"
        if tok != toks[i]:
            break
        lp_ex = OL.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0] if i < len(ex["logits"]) else None
        if i == 0:
            cmp_noise(lp, lp_ref, lp_ex, f"{kind} llava logprobs step 0")
    assert eng.device_error() == 0


def test_vision_feature_cache_through_the_server_loop():
    """the GPU thread fills `VisionFeatureCache` on a miss (models with `encode_image`) and hands a
    hit to the model as `cached_image_features` (server/generation.py:1636-1675, dispatch.py:800-809): the repeat of
    a request produces the same tokens without running the tower again."""
    import types
    from oracle import llava as OL
    from mlx_vlm_b200.server import GenerationArguments, ResponseGenerator
    from mlx_vlm_b200.vision_cache import VisionFeatureCache
    c = OL.LlavaCfg(vision=OL.ClipCfg(hidden_size=64, num_hidden_layers=3, intermediate_size=128,
                                      num_attention_heads=4, image_size=42, patch_size=14),
                    text=OL.LlamaCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512,
                                     num_attention_heads=4, num_key_value_heads=2, vocab_size=320),
                    image_token_index=300)
    W, model = _build(c)
    model.config.eos_token_id = []
    eng = model.engine
    req = OL.synthetic_request(c, n_text=8, seed=1)
    ids = req["input_ids"]
    pv = req["pixel_values"].permute(0, 3, 1, 2).contiguous().cuda()
    proc = types.SimpleNamespace(tokenizer=types.SimpleNamespace(stopping_criteria=None))
    cache = VisionFeatureCache()
    srv = ResponseGenerator(model, proc, vision_cache=cache, decode_slice=2)
    runs, launches = [], []
    for _ in range(3):
        l0 = eng.launch_count
        tower_calls = {"n": 0}
        inner = model.encode_image

        def counted(x, _inner=inner):
            tower_calls["n"] += 1
            return _inner(x)
        model.encode_image = counted
        toks = [e.token for e in srv.generate({"input_ids": ids, "pixel_values": pv}, GenerationArguments(max_tokens=5),
                                              images="image-key-1", timeout=60)]
        model.encode_image = inner
        runs.append(toks)
        launches.append(tower_calls["n"])
    srv.stop_and_join()
    assert srv._error is None, srv._error
    assert launches == [1, 0, 0], launches          # one miss, then hits
    assert len(cache) == 1
    assert runs[0] == runs[1] == runs[2] and len(runs[0]) == 5
    ref = OL.greedy_generate(c, W, ids, req["pixel_values"], 5)
    from oracle.mlx_semantics import Rounder
    for i, tok in enumerate(runs[0]):
        lp_ref = OL.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        assert _token_ok(tok, lp_ref), (i, tok, ref["tokens"][i])
        if tok != ref["tokens"][i]:
            break
