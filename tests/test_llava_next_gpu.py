"""LLaVA-Next product path on the GPU (a sibling of LLaVA-1.5 on the same kernels) against
oracle/llava_next.py: the crops of one image through the fp32-accurate CLIP tower + projector, newline blocks, the
growing merge (image rows pure copies in block order, text rows exact embeddings), greedy generation."""
import numpy as np
import pytest
import torch

from _util import cmp_noise, rl2
from test_engine_gpu import _token_ok

pytestmark = pytest.mark.gpu


def test_llava_next_blocks_merge_and_generate():
    from oracle import llava as OL
    from oracle import llava_next as ON
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.llava_next import Model, ModelConfig, TextConfig, VisionConfig
    c = OL.LlavaCfg(vision=OL.ClipCfg(hidden_size=64, num_hidden_layers=3, intermediate_size=128, num_attention_heads=4,
                                      image_size=42, patch_size=14),
                    text=OL.LlamaCfg(hidden_size=256, num_hidden_layers=2, intermediate_size=512, num_attention_heads=4,
                                     num_key_value_heads=2, vocab_size=320), image_token_index=300)
    W = ON.init_weights(c, 4)
    v, t = c.vision, c.text
    cfg = ModelConfig(text_config=TextConfig(hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                                             intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                                             num_key_value_heads=t.num_key_value_heads, vocab_size=t.vocab_size,
                                             rms_norm_eps=t.rms_norm_eps, rope_theta=t.rope_theta),
                      vision_config=VisionConfig(num_hidden_layers=v.num_hidden_layers, hidden_size=v.hidden_size,
                                                 intermediate_size=v.intermediate_size,
                                                 num_attention_heads=v.num_attention_heads, image_size=v.image_size,
                                                 patch_size=v.patch_size, layer_norm_eps=v.layer_norm_eps),
                      image_token_index=c.image_token_index, vocab_size=t.vocab_size)
    model = Model(cfg, device="cuda:0")
    model.load_weights(W)
    eng = model.engine
    rng = np.random.default_rng(5)
    n_crops, P = 3, v.num_patches
    pv = rng.standard_normal((1, n_crops, 3, v.image_size, v.image_size)).astype(np.float32)
    text = rng.integers(3, 290, size=9).tolist()
    # three <image> tokens: crop 0, crop 1 and crop 2 are inserted; a fourth would take the first newline block
    ids = np.asarray([text[:2] + [300] + text[2:5] + [300] + text[5:7] + [300] + text[7:]])
    n = 4
    ref = ON.greedy_generate(c, W, ids, pv, n)
    ex = ON.greedy_generate(c, W, ids, pv, n, dtype="f32")
    pvd = torch.from_numpy(pv).cuda()
    blocks = model.encode_image(pvd)
    eng.stream.synchronize()
    assert tuple(blocks.shape) == (2 * n_crops, P, t.hidden_size)
    got = blocks.float().cpu()
    assert rl2(got[:n_crops], ref["image_blocks"][:n_crops]) < 1e-3
    assert torch.equal(got[n_crops:], ref["image_blocks"][n_crops:]), "newline blocks are exact copies"
    emb = model.get_input_embeddings(ids, pvd)
    eng.stream.synchronize()
    e_cpu = emb.inputs_embeds[0].float().cpu()
    assert e_cpu.shape[0] == ref["inputs_embeds"].shape[1] == ids.shape[1] - 3 + 3 * P
    from mlx_vlm_b200.models.llava_next.llava_next import merge_plan
    plan, used = merge_plan(ids, 300, 2 * n_crops, P)
    img = [i for i, tk in enumerate(plan) if tk == 300]
    txt = [i for i, tk in enumerate(plan) if tk != 300]
    assert used == 3 and torch.equal(e_cpu[img], got[:3].reshape(-1, t.hidden_size)), "image rows: blocks 0..2 in order"
    assert torch.equal(e_cpu[txt], W["language_model.model.embed_tokens.weight"][torch.tensor([plan[i] for i in txt])])
    for i, (tok, lp) in enumerate(generate_step(ids, model, pvd, None, max_tokens=n)):
        lp_ref = OL.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        assert _token_ok(tok, lp_ref), f"token {i}: {tok} vs {ref['tokens'][i]}"
        if tok != ref["tokens"][i]:
            break
        if i == 0:
            lp_ex = OL.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
            cmp_noise(lp, lp_ref, lp_ex, "llava_next logprobs step 0")
    assert eng.device_error() == 0
