"""Qwen2.5-VL product path on the GPU against oracle/qwen2_5vl.py: the new device ops (2-D rotary on fp32
q / k, attention inside ragged segments, row gather) op by op against plain fp32 references, then the tower
(window-ordered RMSNorm / SwiGLU blocks, windowed + full attention, merger, reverse order), the merge and greedy tokens.
The reference runs this tower in bf16; the product computes it at fp32 accuracy from the bf16-rounded pixels and rounds the
features once, so the end-to-end bar is the noise-relative one (tests/_util.cmp_noise)."""
import numpy as np
import pytest
import torch

from _util import cmp_noise, rl2

pytestmark = pytest.mark.gpu


def _token_ok_noise(tok, lp_ref, lp_ex):
    m = float(lp_ref.max())
    tol = 2 * abs(m) * 2.0 ** -7 + 3 * float((lp_ref.float() - lp_ex.float()).abs().max()) + 1e-6
    return float(lp_ref[tok]) >= m - tol


def _model_cfg(c):
    from mlx_vlm_b200.models.qwen2_5_vl import ModelConfig, TextConfig, VisionConfig
    v, t = c.vision, c.text
    text = TextConfig(model_type="qwen2_5_vl", hidden_size=t.hidden_size, num_hidden_layers=t.num_hidden_layers,
                      intermediate_size=t.intermediate_size, num_attention_heads=t.num_attention_heads,
                      rms_norm_eps=t.rms_norm_eps, vocab_size=t.vocab_size, num_key_value_heads=t.num_key_value_heads,
                      rope_theta=t.rope_theta, rope_scaling={"type": "mrope", "mrope_section": list(t.mrope_section)},
                      tie_word_embeddings=t.tie_word_embeddings)
    vision = VisionConfig(depth=v.depth, hidden_size=v.hidden_size, intermediate_size=v.intermediate_size,
                          out_hidden_size=v.out_hidden_size, num_heads=v.num_heads, patch_size=v.patch_size,
                          in_channels=v.in_channels, spatial_merge_size=v.spatial_merge_size,
                          temporal_patch_size=v.temporal_patch_size, window_size=v.window_size,
                          fullatt_block_indexes=list(v.fullatt_block_indexes))
    return ModelConfig(text_config=text, vision_config=vision, model_type="qwen2_5_vl", image_token_id=c.image_token_id,
                       video_token_id=c.video_token_id, vision_start_token_id=c.vision_start_token_id,
                       vision_end_token_id=c.vision_end_token_id, vocab_size=t.vocab_size)


def _cfg(kind):
    from oracle import qwen2_5vl as O
    if kind == "tiny":
        return O.tiny_cfg()
    # the real tower widths: 1280 / 16 heads (head_dim 80), SwiGLU 3420, windows of 112 px = 4 x 4 merge units
    c = O.tiny_cfg()
    c.vision = O.VisionCfg(depth=3, out_hidden_size=512, fullatt_block_indexes=(1,))
    c.text = O.Q.TextCfg(hidden_size=512, num_hidden_layers=2, intermediate_size=1024, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=1024, mrope_section=(16, 24, 24), tie_word_embeddings=False)
    return c


def _ops():
    from mlx_vlm_b200.models.qwen2_5_vl import Model
    from mlx_vlm_b200.models.tower_ops import TowerOps
    model = Model(_model_cfg(_cfg("tiny")), device="cuda:0")
    return model.engine, TowerOps(model.engine)


def test_rope_varlen_attention_and_gather_ops():
    eng, ops = _ops()
    rng = np.random.default_rng(0)
    for (T, nh, hd, cu) in [(40, 2, 32, [0, 16, 24, 40]), (300, 16, 80, [0, 64, 128, 130, 300]), (70, 3, 64, [0, 70])]:
        E = nh * hd
        qkv = rng.standard_normal((T, 3 * E)).astype(np.float32)
        pos = rng.integers(0, 40, size=(T, 2)).astype(np.int32)
        inv = (1.0 / (10000.0 ** (np.arange(0, hd // 2, 2, dtype=np.float32) / np.float32(hd // 2)))).astype(np.float32)
        with torch.cuda.stream(eng.stream):
            qkv_d, pos_d, inv_d = torch.from_numpy(qkv).cuda(), torch.from_numpy(pos).cuda(), torch.from_numpy(inv).cuda()
            cu_d = torch.tensor(cu, dtype=torch.int32).cuda()
        ops.vision_rope(qkv_d, pos_d, inv_d, nh, hd)
        eng.stream.synchronize()
        # fp32 reference: x * cos + rotate_half(x) * sin with cos / sin of [row freqs | column freqs], tiled twice
        ang = torch.cat([torch.from_numpy(pos[:, :1].astype(np.float32)) * torch.from_numpy(inv)[None],
                         torch.from_numpy(pos[:, 1:].astype(np.float32)) * torch.from_numpy(inv)[None]], 1)
        cos, sin = torch.cos(ang).repeat(1, 2)[:, None], torch.sin(ang).repeat(1, 2)[:, None]
        x = torch.from_numpy(qkv).reshape(T, 3, nh, hd)

        def rot(a):
            return a * cos + torch.cat([-a[..., hd // 2:], a[..., :hd // 2]], -1) * sin
        want = torch.stack([rot(x[:, 0]), rot(x[:, 1]), x[:, 2]], 1).reshape(T, 3 * E)
        got = qkv_d.cpu()
        assert torch.equal(got[:, 2 * E:], want[:, 2 * E:]), "v is untouched"
        assert (got - want).abs().max() < 2e-5, (got - want).abs().max()
        # attention inside ragged segments
        out = ops.f32(T, E)
        ops.attention_varlen((qkv_d, 3 * E, hd), (qkv_d[:, E:], 3 * E, hd), (qkv_d[:, 2 * E:], 3 * E, hd), n_heads=nh, n_kv=nh,
                             hd=hd, cu=cu_d, n_seg=len(cu) - 1, max_len=int(np.diff(cu).max()), scale=hd ** -0.5, out32=out)
        eng.stream.synchronize()
        g = got.reshape(T, 3, nh, hd)
        ref = torch.empty(T, nh, hd)
        for a, b in zip(cu[:-1], cu[1:]):
            q, k, v = (g[a:b, i].transpose(0, 1) for i in range(3))
            ref[a:b] = (torch.softmax(q @ k.transpose(1, 2) * hd ** -0.5, -1) @ v).transpose(0, 1)
        e = rl2(out.cpu(), ref.reshape(T, E))
        assert e < 2e-6, e
    # row gather in units
    x = rng.standard_normal((24, 36)).astype(np.float32)
    idx = rng.permutation(6).astype(np.int32)
    with torch.cuda.stream(eng.stream):
        xd, idd = torch.from_numpy(x).cuda(), torch.from_numpy(idx).cuda()
    out = ops.f32(24, 36)
    ops.gather_rows(xd, idd, 4, out)
    eng.stream.synchronize()
    assert np.array_equal(out.cpu().numpy(), x.reshape(6, 4, 36)[idx].reshape(24, 36))


@pytest.mark.parametrize("kind,grids", [("tiny", [(8, 12)]), ("tiny", [(6, 10), (4, 4)]), ("real_widths", [(20, 12)])])
def test_qwen2_5_vl_tower_merge_and_generate(kind, grids):
    from oracle import qwen2_5vl as O
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200.generate import generate_step
    from mlx_vlm_b200.models.qwen2_5_vl import Model
    c = _cfg(kind)
    W = O.init_weights(c, 0)
    model = Model(_model_cfg(c), device="cuda:0")
    model.load_weights(W)
    eng = model.engine
    rng = np.random.default_rng(2)
    v = c.vision
    K = v.in_channels * v.temporal_patch_size * v.patch_size ** 2
    pv = np.concatenate([rng.standard_normal((h * w, K)).astype(np.float32) for h, w in grids], 0)
    grid = np.asarray([[1, h, w] for h, w in grids], dtype=np.int64)
    text = rng.integers(0, c.image_token_id - 16, size=10).tolist()
    ids = text[:3]
    for h, w in grids:
        ids += [c.vision_start_token_id] + [c.image_token_id] * (h * w // 4) + [c.vision_end_token_id]
    ids = np.asarray([ids + text[3:]], dtype=np.int64)
    n = 4
    ref = O.greedy_generate(c, W, ids, pv, grid, n)
    ex = O.greedy_generate(c, W, ids, pv, grid, n, dtype="f32")
    pvd = torch.from_numpy(pv).cuda()
    feats = model.vision_tower(pvd, grid)
    eng.stream.synchronize()
    assert tuple(feats.shape) == tuple(ref["image_features"].shape)
    cmp_noise(feats.float().cpu(), ref["image_features"], ex["image_features"], f"{kind} {grids} qwen2.5-vl image features")
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    eng.stream.synchronize()
    e_cpu = emb.inputs_embeds[0].float().cpu()
    pos = np.flatnonzero(ids[0] == c.image_token_id)
    assert torch.equal(e_cpu[pos], feats.float().cpu()), "image rows are the tower's features in order"
    rest = np.flatnonzero(ids[0] != c.image_token_id)
    assert torch.equal(e_cpu[rest], W["language_model.model.embed_tokens.weight"][torch.from_numpy(ids[0][rest])])
    qc = O._qcfg(c)
    want_pos, want_delta = O.Q.get_rope_index(qc, ids, grid, None, None)
    assert np.array_equal(np.asarray(emb.position_ids), np.asarray(want_pos))
    for i, (tok, lp) in enumerate(generate_step(ids, model, pvd, None, max_tokens=n, image_grid_thw=grid)):
        lp_ref = O.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        lp_ex = O.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
        if i == 0:
            cmp_noise(lp, lp_ref, lp_ex, f"{kind} qwen2.5-vl logprobs step 0")
        if tok != ref["tokens"][i] or ex["tokens"][i] != ref["tokens"][i]:
            assert _token_ok_noise(tok, lp_ref, lp_ex), f"{kind}: token {i}: {tok} vs {ref['tokens'][i]}"
            break
    assert eng.device_error() == 0


def test_qwen2_5_vl_loads_from_a_checkpoint_directory(tmp_path):
    """HF names (`visual.*`, `model.*`, conv weight [O, C, T, H, W]) + config.json with the text parameters at the root"""
    import json
    import os
    import types
    from safetensors.torch import save_file
    from oracle import qwen2_5vl as O
    from oracle.mlx_semantics import Rounder
    from mlx_vlm_b200 import load
    from mlx_vlm_b200.generate import generate_step
    c = _cfg("tiny")
    W = O.init_weights(c, 3)
    tensors = {}
    for k, x in W.items():
        name = k.replace("vision_tower.", "visual.").replace("language_model.model.", "model.").replace("language_model.", "")
        tensors[name] = x.to(torch.bfloat16).contiguous()
    t, v = c.text, c.vision
    cfg = {"model_type": "qwen2_5_vl", "hidden_size": t.hidden_size, "num_hidden_layers": t.num_hidden_layers,
           "intermediate_size": t.intermediate_size, "num_attention_heads": t.num_attention_heads,
           "num_key_value_heads": t.num_key_value_heads, "rms_norm_eps": t.rms_norm_eps, "vocab_size": t.vocab_size,
           "rope_theta": t.rope_theta, "rope_scaling": {"type": "mrope", "mrope_section": list(t.mrope_section)},
           "tie_word_embeddings": True, "image_token_id": c.image_token_id, "video_token_id": c.video_token_id,
           "vision_start_token_id": c.vision_start_token_id, "vision_end_token_id": c.vision_end_token_id,
           "vision_config": {"model_type": "qwen2_5_vl", "depth": v.depth, "hidden_size": v.hidden_size,
                             "intermediate_size": v.intermediate_size, "out_hidden_size": v.out_hidden_size,
                             "num_heads": v.num_heads, "window_size": v.window_size,
                             "fullatt_block_indexes": list(v.fullatt_block_indexes)}}
    d = str(tmp_path)
    save_file(tensors, os.path.join(d, "model.safetensors"))
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump(cfg, f)
    proc = types.SimpleNamespace(tokenizer=types.SimpleNamespace(stopping_criteria=None))
    model, processor = load(d, processor=proc, device="cuda:0")
    assert type(model).__module__.endswith("models.qwen2_5_vl.qwen2_5_vl")
    req = O.synthetic_request(c, 8, (8, 8), seed=4)
    ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
    ref = O.greedy_generate(c, W, ids, pv, grid, 3)
    ex = O.greedy_generate(c, W, ids, pv, grid, 3, dtype="f32")
    for i, (tok, lp) in enumerate(generate_step(ids, model, torch.from_numpy(pv).cuda(), None, max_tokens=3, image_grid_thw=grid)):
        lp_ref = O.Q.logprobs_from_logits(Rounder("bf16"), ref["logits"][i])[0]
        lp_ex = O.Q.logprobs_from_logits(Rounder("f32"), ex["logits"][i])[0]
        if tok != ref["tokens"][i] or ex["tokens"][i] != ref["tokens"][i]:
            assert _token_ok_noise(tok, lp_ref, lp_ex), f"token {i}: {tok} vs {ref['tokens'][i]}"
            break
    assert model.engine.device_error() == 0
