"""Idefics3 / SmolVLM product path on the GPU against oracle/idefics3.py.  The reference runs this
tower in bf16; the product computes it at fp32 accuracy and rounds where the reference hands the pooled output over, so
the end-to-end bar is the noise-relative one (tests/_util.cmp_noise): no further from the oracle's bf16 result than the
oracle's bf16 result is from the exact evaluation.  The pixel shuffle is a permutation: bit-exact."""
import numpy as np
import pytest
import torch

from _util import cmp_noise, rl2

pytestmark = pytest.mark.gpu


def _token_ok_noise(tok, lp_ref, lp_ex):
    """greedy token acceptable for a tower computed at another precision than the oracle's: its oracle logprob is
    within (2 bf16 ulps + 3x the oracle's own bf16-vs-exact deviation) of the best"""
    m = float(lp_ref.max())
    tol = 2 * abs(m) * 2.0 ** -7 + 3 * float((lp_ref.float() - lp_ex.float()).abs().max()) + 1e-6
    return float(lp_ref[tok]) >= m - tol


def _model_cfg(c, smol=False):
    if smol:
        from mlx_vlm_b200.models.smolvlm import ModelConfig, TextConfig, VisionConfig
    else:
        from mlx_vlm_b200.models.idefics3 import ModelConfig, TextConfig, VisionConfig
    v, t = c.vision, c.text
    return ModelConfig(
        text_config=TextConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                               intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                               num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                               rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta),
        vision_config=VisionConfig(hidden_size=v.hidden_size, num_hidden_layers=v.num_hidden_layers,
                                   intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                                   image_size=v.image_size, patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps),
        scale_factor=c.scale_factor, image_token_id=c.image_token_index, vocab_size=t.vocab_size)


def _cfg(kind):
    from oracle import idefics3 as O3
    I2 = O3.I2
    if kind == "tiny":
        return O3.Idefics3Cfg(
            vision=I2.SiglipCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=96, num_attention_heads=4,
                                image_size=84, patch_size=14),
            text=I2.MistralCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                               num_key_value_heads=2, vocab_size=320), image_token_index=300)
    if kind == "smolvlm_widths":    # SmolVLM2-2.2B tower widths: 1152 / 18 heads (head_dim 64), mlp 4304; scale 4
        return O3.Idefics3Cfg(
            vision=I2.SiglipCfg(num_hidden_layers=2, num_attention_heads=18, image_size=384),
            text=I2.MistralCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024, num_attention_heads=4,
                               num_key_value_heads=2, vocab_size=49155, rope_theta=130000.0),
            scale_factor=4, image_token_index=49153)
    # Idefics3-8B tower: SigLIP-SO400M widths (16 heads of 72), 364 px grid, scale 2
    return O3.Idefics3Cfg(
        vision=I2.SiglipCfg(num_hidden_layers=2, image_size=364),
        text=I2.MistralCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024, num_attention_heads=4,
                           num_key_value_heads=2, vocab_size=128259, rope_theta=500000.0), image_token_index=128257)


def test_pixel_shuffle_kernel_is_the_reference_permutation():
    import json, os
    from mlx_vlm_b200.engine import Engine
    from mlx_vlm_b200.models.idefics3 import Model
    from mlx_vlm_b200.models.tower_ops import SplitBuf, TowerOps
    from oracle import idefics3 as O3
    model = Model(_model_cfg(_cfg("tiny")), device="cuda:0")
    eng = model.engine
    ops = TowerOps(eng)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "idefics3_golden.json")) as f:
        G = json.load(f)
    rng = np.random.default_rng(0)
    for n_img, side, E, s in [(1, 4, 8, 2), (3, 6, 12, 3), (2, 26, 1152, 2), (1, 24, 768, 4)]:
        x = rng.standard_normal((n_img, side * side, E)).astype(np.float32)
        want = O3.pixel_shuffle(torch.from_numpy(x), s).reshape(-1, E * s * s)
        with torch.cuda.stream(eng.stream):
            xd = torch.from_numpy(x).cuda().reshape(-1, E)
        for rnd in (False, True):
            out = SplitBuf(eng, want.shape[0], E * s * s)
            ops.pixel_shuffle(xd, n_img, side, s, out, round_in=rnd)
            eng.stream.synchronize()
            hi = out.t[:, :E * s * s].float().cpu()
            lo = out.t[:, out.n_pad:out.n_pad + E * s * s].float().cpu()
            if rnd:
                assert torch.equal(hi, want.to(torch.bfloat16).float()) and not lo.any()
            else:
                assert torch.equal(hi, want.to(torch.bfloat16).float())
                assert (hi + lo - want).abs().max() <= want.abs().max() * 2.0 ** -16
            assert not out.t[:, E * s * s:out.n_pad].any()
    # the reference's own small cases
    for c in G["pixel_shuffle"]:
        seq, E = c["side"] ** 2, c["E"]
        if E % 4:
            continue
        x = (np.arange(seq, dtype=np.float32)[:, None] * 100 + np.arange(E, dtype=np.float32)[None, :])
        xd = torch.from_numpy(x).cuda()
        out = SplitBuf(eng, c["out_shape"][1], c["out_shape"][2])
        ops.pixel_shuffle(xd, 1, c["side"], c["scale"], out, round_in=False)
        eng.stream.synchronize()
        got = (out.t[:, :out.n].float() + out.t[:, out.n_pad:out.n_pad + out.n].float()).cpu().numpy()
        assert np.array_equal(got, np.asarray(c["out"], dtype=np.float32))


@pytest.mark.parametrize("kind", ["tiny", "smolvlm_widths", "idefics3_widths"])
def test_idefics3_features_merge_and_generate(kind):
    from oracle import idefics3 as O3
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200.generate import generate_step
    c = _cfg(kind)
    smol = kind == "smolvlm_widths"
    if smol:
        from mlx_vlm_b200.models.smolvlm import Model
    else:
        from mlx_vlm_b200.models.idefics3 import Model
    W = O3.init_weights(c, 0)
    model = Model(_model_cfg(c, smol), device="cuda:0")
    model.load_weights(W)
    eng = model.engine
    rng = np.random.default_rng(5)
    side = {"tiny": 84, "smolvlm_widths": 112, "idefics3_widths": 112}[kind]
    per_img = (side // 14 // c.scale_factor) ** 2
    # 3 image slots: slot 1 is an all-zero padding image; slot 2 is only partly valid (ragged pixel mask)
    pv = rng.standard_normal((1, 3, 3, side, side)).astype(np.float32)
    pv[0, 1] = 0.0
    pam = np.ones((1, 3, side, side), dtype=bool)
    pam[0, 2, side - 28:, :] = False
    pam[0, 2, :, side - 14:] = False
    text = rng.integers(3, min(c.image_token_index, 30000) - 1, size=8).tolist()
    ids = np.asarray([text[:4] + [c.image_token_index] * (2 * per_img) + text[4:]])
    n = 4
    ref = O3.greedy_generate(c, W, ids, pv, pam, n)
    ex = O3.greedy_generate(c, W, ids, pv, pam, n, dtype="f32")
    pv_dev = torch.from_numpy(pv).cuda()
    feats = model.encode_image(pv_dev, pam)
    eng.stream.synchronize()
    assert tuple(feats.shape) == (2 * per_img, c.text.hidden_size)
    cmp_noise(feats.float().cpu(), ref["image_features"].reshape(-1, c.text.hidden_size),
              ex["image_features"].reshape(-1, c.text.hidden_size), f"{kind} idefics3 image features")
    emb = model.get_input_embeddings(ids, pv_dev, pixel_attention_mask=pam)
    eng.stream.synchronize()
    e_cpu = emb.inputs_embeds[0].float().cpu()
    pos = np.flatnonzero(ids[0] == c.image_token_index)
    assert torch.equal(e_cpu[pos], feats.float().cpu()), "image rows are pure copies in feature order"
    rest = np.flatnonzero(ids[0] != c.image_token_index)
    assert torch.equal(e_cpu[rest], W["language_model.embed_tokens.weight"][torch.from_numpy(ids[0][rest])])
    with pytest.raises(ValueError, match="do not match"):
        model._prepare_inputs_for_multimodal(feats[:-1], None, ids)
    for i, (tok, lp) in enumerate(generate_step(ids, model, pv_dev, None, max_tokens=n, pixel_attention_mask=pam)):
        lp_ref = O3.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        lp_ex = O3.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
        if i == 0:
            cmp_noise(lp, lp_ref, lp_ex, f"{kind} idefics3 logprobs step 0")
        if tok != ref["tokens"][i] or ex["tokens"][i] != ref["tokens"][i]:
            assert _token_ok_noise(tok, lp_ref, lp_ex), f"{kind}: token {i}: {tok} vs {ref['tokens'][i]}"
            break   # a near-tie: histories legitimately diverge from here
    assert eng.device_error() == 0


def test_smolvlm_loads_from_a_checkpoint_directory(tmp_path):
    """`load()` on a directory with config.json (model_type smolvlm) + safetensors under the HF names (model.* prefixes,
    lm_head at the top level, PyTorch conv layout) -> models/smolvlm, and generates like the oracle"""
    import json
    import os
    import types
    from safetensors.torch import save_file
    from oracle import idefics3 as O3
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200 import load
    from mlx_vlm_b200.generate import generate_step
    c = _cfg("tiny")
    W = O3.init_weights(c, 7)
    tensors = {}
    for k, x in W.items():
        if "patch_embedding.weight" in k:
            x = x.permute(0, 3, 1, 2)
        if k.startswith("language_model.lm_head"):
            name = "lm_head." + k.split("lm_head.", 1)[1]
        elif k.startswith("language_model."):
            name = "model.text_model." + k[len("language_model."):]
        else:
            name = "model." + k
        tensors[name] = x.to(torch.bfloat16).contiguous()
    v, t = c.vision, c.text
    cfg = {"model_type": "smolvlm", "image_token_id": c.image_token_index, "vocab_size": t.vocab_size, "scale_factor": 2,
           "text_config": {"model_type": "llama", "hidden_size": t.hidden_size, "num_hidden_layers": t.num_hidden_layers,
                           "intermediate_size": t.intermediate_size, "num_attention_heads": t.num_attention_heads,
                           "num_key_value_heads": t.num_key_value_heads, "vocab_size": t.vocab_size,
                           "rms_norm_eps": t.rms_norm_eps, "rope_theta": t.rope_theta},
           "vision_config": {"hidden_size": v.hidden_size, "num_hidden_layers": v.num_hidden_layers,
                             "intermediate_size": v.intermediate_size, "num_attention_heads": v.num_attention_heads,
                             "image_size": v.image_size, "patch_size": v.patch_size, "layer_norm_eps": v.layer_norm_eps}}
    d = str(tmp_path)
    save_file(tensors, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    proc = types.SimpleNamespace(tokenizer=types.SimpleNamespace(stopping_criteria=None))
    model, processor = load(d, processor=proc, device="cuda:0")
    assert type(model).__module__.endswith("models.smolvlm") and processor is proc
    rng = np.random.default_rng(1)
    pv = rng.standard_normal((1, 1, 3, 84, 84)).astype(np.float32)
    ids = np.asarray([[5, 9] + [c.image_token_index] * 9 + [11, 12, 13]])
    ref = O3.greedy_generate(c, W, ids, pv, None, 3)
    ex = O3.greedy_generate(c, W, ids, pv, None, 3, dtype="f32")
    for i, (tok, lp) in enumerate(generate_step(ids, model, torch.from_numpy(pv).cuda(), None, max_tokens=3)):
        lp_ref = O3.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        lp_ex = O3.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
        if tok != ref["tokens"][i] or ex["tokens"][i] != ref["tokens"][i]:
            assert _token_ok_noise(tok, lp_ref, lp_ex), f"token {i}: {tok} vs {ref['tokens'][i]}"
            break
    assert model.engine.device_error() == 0
