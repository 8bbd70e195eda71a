"""Idefics2 product path on the GPU against oracle/idefics2.py: SigLIP tower with
bucketed position ids + modality projection + Perceiver resampler in fp32 with bf16-valued weights (the
reference never casts float32 pixel_values), ONE rounding at the merge, Mistral LM in bf16."""
import numpy as np
import pytest
import torch

from _util import cmp_noise, rl2
from test_engine_gpu import _token_ok

pytestmark = pytest.mark.gpu


def _model_cfg(c):
    from mlx_vlm_b200.models.idefics2.config import ModelConfig, PerceiverConfig, TextConfig, VisionConfig
    v, t, p = c.vision, c.text, c.perceiver
    return ModelConfig(
        text_config=TextConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                               intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                               num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                               rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta),
        vision_config=VisionConfig(hidden_size=v.hidden_size, num_hidden_layers=v.num_hidden_layers,
                                   intermediate_size=v.intermediate_size, num_attention_heads=v.num_attention_heads,
                                   image_size=v.image_size, patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps),
        perceiver_config=PerceiverConfig(num_key_value_heads=p.num_key_value_heads, resampler_depth=p.resampler_depth,
                                         resampler_head_dim=p.resampler_head_dim, resampler_n_heads=p.resampler_n_heads,
                                         resampler_n_latents=p.resampler_n_latents),
        image_token_id=c.image_token_index, vocab_size=t.vocab_size)


def _cfg(kind):
    from oracle import idefics2 as OI
    if kind == "tiny":       # decoder engine head_dim 64; tower head_dim 16; perceiver head_dim 16, GQA 4:2
        return OI.Idefics2Cfg(
            vision=OI.SiglipCfg(hidden_size=64, num_hidden_layers=2, intermediate_size=96, num_attention_heads=4,
                                image_size=70, patch_size=14),
            text=OI.MistralCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                               num_key_value_heads=2, vocab_size=320),
            perceiver=OI.PerceiverCfg(num_key_value_heads=2, resampler_depth=2, resampler_head_dim=16,
                                      resampler_n_heads=4, resampler_n_latents=6),
            image_token_index=300)
    if kind == "siglip_full_depth":   # all 27 SigLIP-SO400M layers + the 3-layer perceiver of Idefics2-8B
        return OI.Idefics2Cfg(
            vision=OI.SiglipCfg(image_size=980),
            text=OI.MistralCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024, num_attention_heads=4,
                               num_key_value_heads=2, vocab_size=32003),
            perceiver=OI.PerceiverCfg(), image_token_index=32001)
    # SigLIP-SO400M widths (1152 / 16 heads = head_dim 72, mlp 4304), 2 layers, 154 px images (11 x 11 patches on
    # a 70 x 70 position grid); the real perceiver geometry (16 heads of 96 over 4 kv heads, 64 latents)
    return OI.Idefics2Cfg(
        vision=OI.SiglipCfg(num_hidden_layers=2, image_size=980),
        text=OI.MistralCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024, num_attention_heads=4,
                           num_key_value_heads=2, vocab_size=32003),
        perceiver=OI.PerceiverCfg(resampler_depth=2), image_token_index=32001)


@pytest.mark.parametrize("kind", ["tiny", "siglip_widths", "siglip_full_depth"])
def test_idefics2_features_merge_and_generate(kind):
    from oracle import idefics2 as OI
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.idefics2 import Model
    c = _cfg(kind)
    W = OI.init_weights(c, 0)
    model = Model(_model_cfg(c), device="cuda:0")
    model.load_weights(W)
    eng = model.engine
    rng = np.random.default_rng(3)
    side = 70 if kind == "tiny" else 154
    n_lat = c.perceiver.resampler_n_latents
    # 3 image slots: slot 1 is an all-zero padding image; slot 2 is only partly valid (ragged pixel mask)
    pv = rng.standard_normal((1, 3, 3, side, side)).astype(np.float32)
    pv[0, 1] = 0.0
    pam = np.ones((1, 3, side, side), dtype=bool)
    pam[0, 2, side - 28:, :] = False
    pam[0, 2, :, side - 14:] = False
    text = rng.integers(3, c.image_token_index - 1, size=8).tolist()
    ids = np.asarray([text[:4] + [c.image_token_index] * (2 * n_lat) + text[4:]])
    n = 4
    ref = OI.greedy_generate(c, W, ids, pv, pam, n)
    ex = OI.greedy_generate(c, W, ids, pv, pam, n, dtype="f32")
    pv_dev = torch.from_numpy(pv).cuda()
    feats = model.encode_image(pv_dev, pam)
    eng.stream.synchronize()
    want32 = OI.image_features(c, W, pv, pam, Rounder("f32")).reshape(-1, c.text.hidden_size)
    got = feats.float().cpu()
    e = rl2(got, want32)
    want_bf = ref["image_features"].reshape(-1, c.text.hidden_size)
    mism = float((got.reshape(-1) != want_bf.reshape(-1)).float().mean())
    print(f"{kind}: connector output vs the fp32 reference path rel={e:.2e}; bf16 elements off the oracle's rounding: {mism:.4f}")
    assert tuple(feats.shape) == (2 * n_lat, c.text.hidden_size)
    assert rl2(got, want_bf) < 1e-3
    assert mism < 0.02
    emb = model.get_input_embeddings(ids, pv_dev, pixel_attention_mask=pam)
    eng.stream.synchronize()
    e_cpu = emb.inputs_embeds[0].float().cpu()
    pos = np.flatnonzero(ids[0] == c.image_token_index)
    assert torch.equal(e_cpu[pos], got), "image rows are pure copies in feature order"
    rest = np.flatnonzero(ids[0] != c.image_token_index)
    assert torch.equal(e_cpu[rest], W["language_model.embed_tokens.weight"][torch.from_numpy(ids[0][rest])])
    with pytest.raises(ValueError, match="do not match"):
        model._prepare_inputs_for_multimodal(feats[:-1], None, ids)
    for i, (tok, lp) in enumerate(generate_step(ids, model, pv_dev, None, max_tokens=n, pixel_attention_mask=pam)):
        lp_ref = OI.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        assert _token_ok(tok, lp_ref), f"{kind}: token {i}: {tok} vs {ref['tokens'][i]}"
        if tok != ref["tokens"][i]:
            break
        if i == 0:
            lp_ex = OI.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
            cmp_noise(lp, lp_ref, lp_ex, f"{kind} idefics2 logprobs step 0")
    assert eng.device_error() == 0
