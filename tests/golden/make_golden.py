"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN FUNCTION SOURCE.

`mlx` (the array library the reference runs on) is not installable offline, and
importing the `mlx_vlm` package pulls in Metal-kernel modules.  The functions on
the hot path that are pure array bookkeeping, however, only use a dozen array ops
that have exact numpy equivalents.  This script therefore

  1. parses the reference files under /root/reference with `ast`,
  2. extracts the UNMODIFIED source of the functions / methods listed in TARGETS,
  3. executes them against a minimal numpy-backed stand-in for `mlx.core`
     (`mx.arange = np.arange`, ... — no arithmetic is re-implemented),
  4. runs them on seeded inputs and writes inputs + outputs to
     tests/golden/reference_golden.json.

Run it in the build container only (it reads /root/reference; the GPU box does
not have it):   python tests/golden/make_golden.py
The CPU test-suite (tests/test_oracle_golden.py) checks the oracle AND the product's
host code against the committed JSON.
"""
import ast
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference/mlx_vlm"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")


# ---------------------------------------------------------------- mx stand-in
def make_mx():
    mx = types.ModuleType("mx_numpy_standin")
    for name in ("arange", "broadcast_to", "ones", "ones_like", "zeros", "zeros_like", "stack",
                 "concatenate", "expand_dims", "where", "cumsum", "sum", "repeat", "tile",
                 "transpose", "max", "outer", "cos", "sin", "take", "maximum", "minimum", "pad"):
        setattr(mx, name, getattr(np, name))
    mx.array = lambda x, dtype=None: np.array(x, dtype=dtype)
    mx.int32, mx.int64, mx.float32, mx.bool_ = np.int32, np.int64, np.float32, np.bool_
    mx.eval = lambda *a, **k: None
    mx.contiguous = np.ascontiguousarray
    # ops used by the samplers / BatchKVCache (still pure numpy, nothing re-implemented)
    for name in ("argsort", "take_along_axis", "exp", "abs", "argpartition", "std", "log"):
        setattr(mx, name, getattr(np, name))
    mx.max = lambda a, axis=None, keepdims=False: np.max(a, axis=axis, keepdims=keepdims)
    mx.sum = lambda a, axis=None, keepdims=False: np.sum(a, axis=axis, keepdims=keepdims)
    mx.std = lambda a, axis=None, keepdims=False: np.std(a, axis=axis, keepdims=keepdims)
    mx.inf = np.inf

    def put_along_axis(a, idx, vals, axis):  # mlx returns a new array
        out = np.array(a, copy=True)
        np.put_along_axis(out, idx, vals, axis=axis)
        return out
    mx.put_along_axis = put_along_axis

    def softmax(x, axis=-1):
        z = np.exp(x - np.max(x, axis=axis, keepdims=True))
        return z / np.sum(z, axis=axis, keepdims=True)
    mx.softmax = softmax

    def compile_(fn=None, **kw):
        return fn if fn is not None else (lambda f: f)
    mx.compile = compile_
    return mx


def extract(path, name, cls=None):
    """Return the source text of function `name` (optionally a method of `cls`)."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name == name)
    seg = ast.get_source_segment(src, node)
    # strip decorators (mx.compile etc. are identity here); dedent methods
    lines = src.splitlines()[node.lineno - 1:node.end_lineno]
    import textwrap
    return textwrap.dedent("\n".join(lines)), f"{path}:{node.lineno}-{node.end_lineno}"


def load(ns, path, name, cls=None):
    code, where = extract(path, name, cls)
    exec(compile(code, f"<reference {where}>", "exec"), ns)
    return ns[name], where


def tolist(x):
    return np.asarray(x).tolist()


def main():
    mx = make_mx()
    from typing import Optional, Sequence
    ns = {"mx": mx, "np": np, "Optional": Optional, "Sequence": Sequence}
    provenance = {}
    golden = {"_about": "outputs of the reference's own function source executed over a "
                        "numpy stand-in for mlx.core (tests/golden/make_golden.py)",
              "provenance": provenance}

    # ---------------- get_rope_index (language.py:216-402)
    get_rope_index, w = load(ns, "models/qwen2_vl/language.py", "get_rope_index", "LanguageModel")
    provenance["get_rope_index"] = w
    stub = types.SimpleNamespace(config=types.SimpleNamespace(
        vision_config=types.SimpleNamespace(spatial_merge_size=2),
        image_token_id=100, video_token_id=101, vision_start_token_id=99))
    rng = np.random.default_rng(0)
    cases = []

    def run_rope(ids, igrid=None, vgrid=None, mask=None, tag=""):
        a = mx.array(ids)
        pos, delta = get_rope_index(stub, a,
                                    None if igrid is None else mx.array(igrid),
                                    None if vgrid is None else mx.array(vgrid),
                                    None if mask is None else mx.array(mask))
        cases.append({"tag": tag, "input_ids": tolist(ids), "image_grid_thw": igrid,
                      "video_grid_thw": vgrid, "attention_mask": mask,
                      "position_ids": tolist(pos), "rope_deltas": tolist(delta)})

    two = [1, 2, 99, 100, 100, 100, 100, 5, 99, 100, 100, 100, 100, 7]
    run_rope([two], [[1, 4, 4], [1, 4, 4]], tag="reference test: two images")
    run_rope([[0, 0] + two], [[1, 4, 4], [1, 4, 4]], mask=[[0, 0] + [1] * len(two)],
             tag="reference test: left padding")
    run_rope([[0, 0, 0, 0], [10, 99, 100, 11]], [[1, 2, 2]], mask=[[0, 0, 0, 0], [1, 1, 1, 1]],
             tag="reference test: fully masked row")
    # one 336x336 image (grid 24x24 -> 144 tokens) inside a 32-token text prompt
    ids = rng.integers(0, 90, size=8).tolist() + [99] + [100] * 144 + [98] + rng.integers(0, 90, size=22).tolist()
    run_rope([ids], [[1, 24, 24]], tag="C1 layout: 336x336 image")
    # video block (t=2) then an image, batch of 2 with different lengths (left padded)
    r0 = [3, 99] + [101] * (2 * 2 * 3) + [4, 5, 99] + [100] * 4 + [6]
    r1 = [7, 8, 9, 99] + [100] * 6 + [10]
    L = max(len(r0), len(r1))
    pad = lambda r: [0] * (L - len(r)) + r
    m = lambda r: [0] * (L - len(r)) + [1] * len(r)
    run_rope([pad(r0), pad(r1)], [[1, 4, 4], [1, 4, 6]], [[2, 4, 6]], mask=[m(r0), m(r1)],
             tag="video + images, batch 2, left padded")
    run_rope([rng.integers(0, 90, size=11).tolist()], tag="text only, no mask")
    run_rope([[0, 0, 0, 5, 6, 7, 8], [1, 2, 3, 4, 5, 6, 7]],
             mask=[[0, 0, 0, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1]], tag="text only, with mask")
    golden["get_rope_index"] = {"config": {"spatial_merge_size": 2, "image_token_id": 100,
                                           "video_token_id": 101, "vision_start_token_id": 99},
                                "cases": cases}

    # ---------------- merge_input_ids_with_image_features (qwen2_vl.py:78-148)
    merge, w = load(ns, "models/qwen2_vl/qwen2_vl.py", "merge_input_ids_with_image_features", "Model")
    provenance["merge_input_ids_with_image_features"] = w
    mcases = []

    def run_merge(ids, n_feats, tag, image_id=100, video_id=101):
        ids = np.asarray(ids)
        B, T = ids.shape
        H = 4
        feats = (1000 + np.arange(n_feats * H, dtype=np.float32)).reshape(n_feats, H)
        emb = -(np.arange(B * T * H, dtype=np.float32) + 1).reshape(B, T, H)
        try:
            out = merge(image_id, video_id, mx.array(feats), mx.array(emb), mx.array(ids))
            out = tolist(out)
            err = None
        except ValueError as e:
            out, err = None, str(e)
        mcases.append({"tag": tag, "input_ids": ids.tolist(), "n_feats": n_feats, "hidden": H,
                       "output": out, "error": err})

    run_merge([[5, 100, 100, 100, 6, 7]], 3, "one image")
    run_merge([[5, 100, 100, 6, 100, 100, 100, 7], [100, 1, 2, 3, 100, 100, 4, 8]], 8, "two rows, interleaved")
    run_merge([[5, 101, 101, 6]], 2, "video fallback (no image token)")
    run_merge([[5, 6, 7]], 2, "no vision tokens at all")
    run_merge([[5, 100, 100, 100, 6]], 2, "too few features -> ValueError")
    run_merge([[5, 100, 100, 6]], 5, "more features than positions (extra ignored)")
    golden["merge_input_ids_with_image_features"] = mcases

    # ---------------- vision rot_pos_emb ids + freqs (vision.py:53-65, 219-255)
    ns["nn"] = types.SimpleNamespace(Module=object)
    load(ns, "models/qwen2_vl/vision.py", "VisionRotaryEmbedding")
    rot_pos_emb, w2 = load(ns, "models/qwen2_vl/vision.py", "rot_pos_emb", "VisionModel")
    provenance["rot_pos_emb"] = w2
    vstub = types.SimpleNamespace(spatial_merge_size=2, rotary_pos_emb=ns["VisionRotaryEmbedding"](40))
    vcases = []
    for grid in ([[1, 4, 6]], [[1, 24, 24]], [[2, 4, 4], [1, 6, 2]]):
        fr = rot_pos_emb(vstub, mx.array(grid))
        vcases.append({"grid_thw": grid, "freqs_shape": list(fr.shape),
                       "freqs_sum": float(np.asarray(fr, dtype=np.float64).sum()),
                       "freqs_head": tolist(np.asarray(fr)[:6, :]),
                       "freqs_tail": tolist(np.asarray(fr)[-3:, :])})
    golden["rot_pos_emb"] = vcases

    # ---------------- apply_rotary_pos_emb_vision / rotate_half (vision.py:28-50), fp32
    load(ns, "models/qwen2_vl/vision.py", "rotate_half")
    arv, w = load(ns, "models/qwen2_vl/vision.py", "apply_rotary_pos_emb_vision")
    provenance["apply_rotary_pos_emb_vision"] = w
    x = rng.standard_normal((1, 5, 2, 8)).astype(np.float32)
    fr = rng.standard_normal((5, 4)).astype(np.float32)
    golden["apply_rotary_pos_emb_vision"] = {"x": tolist(x), "freqs": tolist(fr),
                                             "out": tolist(arv(mx.array(x), mx.array(fr)))}

    # ---------------- M-RoPE pieces (rope_utils.py:519-526, 1289-1334), fp32
    sel, w = load(ns, "models/rope_utils.py", "_chunked_position_selector")
    provenance["_chunked_position_selector"] = w
    golden["chunked_position_selector"] = {"mrope_section": [16, 24, 24], "freq_dim": 64,
                                           "selector": tolist(sel([16, 24, 24], 64)),
                                           "small": tolist(sel([2, 3, 3], 8))}
    load(ns, "models/rope_utils.py", "rotate_half")
    are, w = load(ns, "models/rope_utils.py", "_apply_rotary_embedding")
    provenance["_apply_rotary_embedding"] = w
    q = rng.standard_normal((1, 2, 3, 8)).astype(np.float32)
    k = rng.standard_normal((1, 1, 3, 8)).astype(np.float32)
    ang = rng.standard_normal((1, 1, 3, 8)).astype(np.float32)
    qe, ke = are(mx.array(q), mx.array(k), np.cos(ang), np.sin(ang), ns["rotate_half"])
    golden["apply_rotary_embedding"] = {"q": tolist(q), "k": tolist(k), "angles": tolist(ang),
                                        "q_out": tolist(qe), "k_out": tolist(ke)}

    # ---------------- KVCache bookkeeping (cache.py:337-439) + create_causal_mask (:24-42)
    ccm, w = load(ns, "models/cache.py", "create_causal_mask")
    provenance["create_causal_mask"] = w
    golden["create_causal_mask"] = [{"N": n, "offset": o, "mask": tolist(ccm(n, o).astype(np.int32))}
                                    for n, o in ((1, 0), (4, 0), (3, 5))]
    ns["_BaseCache"] = object
    ns["BatchKVCache"] = None
    ns["QuantizedKVCache"] = None
    ns["create_attention_mask"] = lambda *a, **k: None
    ns["tree_reduce"] = None
    code, w = extract("models/cache.py", "KVCache")
    exec(compile(code, "<ref KVCache>", "exec"), ns)
    provenance["KVCache"] = w
    c = ns["KVCache"]()
    trace = []
    for L in (5, 1, 1, 300, 1):
        kk = rng.standard_normal((1, 2, L, 4)).astype(np.float32)
        ks, vs = c.update_and_fetch(mx.array(kk), mx.array(kk + 1))
        trace.append({"L": L, "offset": int(c.offset), "capacity": int(c.keys.shape[2]),
                      "returned_len": int(ks.shape[2]),
                      "last_key_sum": float(np.asarray(ks)[..., -1, :].sum())})
    n = c.trim(7)
    trace.append({"trim": 7, "trimmed": int(n), "offset": int(c.offset),
                  "state_len": int(c.state[0].shape[2])})
    golden["KVCache_trace"] = trace

    # ---------------- BatchKVCache bookkeeping (cache.py:972-1201) + prompt padding (ar.py:548-560)
    ns["List"] = list
    code, w = extract("models/cache.py", "dynamic_roll")
    exec(compile(code, "<ref dynamic_roll>", "exec"), ns)
    ns["_BaseCache"] = object
    code, w = extract("models/cache.py", "BatchKVCache")
    exec(compile(code, "<ref BatchKVCache>", "exec"), ns)
    provenance["BatchKVCache"] = w
    KV, BKV = ns["KVCache"], ns["BatchKVCache"]

    def filled(n, base):  # a KVCache holding n positions with recognisable values
        c = KV()
        if n:
            kk = (base + np.arange(n, dtype=np.float32)).reshape(1, 1, n, 1) * np.ones((1, 2, 1, 2), np.float32)
            c.update_and_fetch(mx.array(kk), mx.array(-kk))
        return c

    def snap(c, tag):
        k = None if c.keys is None else tolist(np.asarray(c.keys)[..., : c._idx, :])
        v = None if c.values is None else tolist(np.asarray(c.values)[..., : c._idx, :])
        return {"op": tag, "left_padding": tolist(c.left_padding), "offset": tolist(c.offset),
                "idx": int(c._idx), "keys": k, "values": v}

    btrace = []
    b1 = BKV.merge([filled(5, 100), filled(2, 200), filled(7, 300)])
    btrace.append(snap(b1, "merge lengths 5,2,7"))
    step = (900 + np.arange(3, dtype=np.float32)).reshape(3, 1, 1, 1) * np.ones((1, 2, 1, 2), np.float32)
    b1.update_and_fetch(mx.array(step), mx.array(-step))
    btrace.append(snap(b1, "append one position"))
    ex = b1.extract(1)
    btrace.append({"op": "extract row 1", "offset": int(ex.offset), "keys": tolist(np.asarray(ex.keys)[..., : ex.offset, :])})
    b1.filter(mx.array([0, 1]))
    btrace.append(snap(b1, "filter keep rows 0,1 (drops the longest: shifts left)"))
    b2 = BKV.merge([filled(3, 400)])
    b1.extend(b2)
    btrace.append(snap(b1, "extend with a 1-row cache of length 3"))
    n = b1.trim(2)
    btrace.append(snap(b1, f"trim 2 (trimmed {int(n)})"))
    b3 = BKV([1, 3, 0])
    btrace.append(snap(b3, "constructor left_padding [1,3,0]"))
    golden["BatchKVCache_trace"] = btrace
    golden["causal_mask_left_padding"] = [
        {"N": n_, "offset": o, "left_padding": lp,
         "mask": tolist(ccm(n_, o, left_padding=mx.array(lp)).astype(np.int32))}
        for n_, o, lp in ((3, 0, [0, 2]), (1, 4, [1, 0, 3]))]
    lp_fn, w = load(ns, "generate/ar.py", "_left_pad_prompts")
    rp_fn, _ = load(ns, "generate/ar.py", "_right_pad_prompts")
    provenance["_left_pad_prompts"] = w
    prompts = [[1, 3, 5], [7], [2, 6, 8, 9]]
    golden["pad_prompts"] = {"prompts": prompts, "left": tolist(lp_fn(prompts)), "right": tolist(rp_fn(prompts)),
                             "left_max6": tolist(lp_fn(prompts, 6))}

    # ---------------- LLaVA merge (models/llava/llava.py:90-116): positional assignment, batch 1
    lmerge, w = load(ns, "models/llava/llava.py", "_merge_input_ids_with_image_features", "Model")
    provenance["llava._merge_input_ids_with_image_features"] = w
    lstub = types.SimpleNamespace(config=types.SimpleNamespace(image_token_index=100))
    lcases = []

    def run_lmerge(ids, n_feats, tag):
        ids = np.asarray(ids)
        H = 4
        feats = (1000 + np.arange(n_feats * H, dtype=np.float32)).reshape(1, n_feats, H)
        emb = -(np.arange(ids.size * H, dtype=np.float32) + 1).reshape(1, ids.shape[1], H)
        try:
            out = tolist(lmerge(lstub, mx.array(feats), mx.array(emb), mx.array(ids)))
            err = None
        except Exception as e:  # ValueError (too many features) or numpy's shape error (too few)
            out, err = None, f"{type(e).__name__}: {e}"
        lcases.append({"tag": tag, "input_ids": ids.tolist(), "n_feats": n_feats, "hidden": H,
                       "positions": np.where(ids == 100)[1].tolist(), "output": out, "error": err})

    run_lmerge([[5, 100, 100, 100, 6, 7]], 3, "one image, three positions")
    run_lmerge([[100, 5, 100, 6, 100, 100]], 4, "scattered positions")
    run_lmerge([[5, 6, 7]], 0, "no image tokens, no features")
    run_lmerge([[5, 100, 100, 6]], 3, "more features than positions -> ValueError")
    run_lmerge([[5, 100, 100, 100, 6]], 2, "fewer features than positions -> shape error")
    golden["llava_merge"] = lcases

    # ---------------- Idefics2 integer logic (models/idefics2/vision.py:123-173, idefics2.py:15-33,185-280)
    mx.flatten = lambda a, start_axis=0, end_axis=-1: np.reshape(
        a, a.shape[:start_axis] + (-1,) + (a.shape[end_axis + 1:] if end_axis != -1 and end_axis + 1 < a.ndim else ()))
    mx.reshape = np.reshape
    mx.uint32 = np.uint32
    code, w = extract("models/idefics2/vision.py", "VisionEmbeddings")
    ns["VisionConfig"] = object
    exec(compile(code, "<ref idefics2 VisionEmbeddings>", "exec"), ns)
    provenance["idefics2.VisionEmbeddings.__call__"] = w
    seen = {}

    def make_embed_stub(num_side, patch):
        st = types.SimpleNamespace(patch_size=patch, num_patches=num_side)
        st.patch_embedding = lambda x: np.zeros((x.shape[0], x.shape[1] // patch, x.shape[2] // patch, 2), np.float32)

        def pos_emb(ids):
            seen["ids"] = np.asarray(ids).copy()
            return np.zeros(ids.shape + (2,), np.float32)
        st.position_embedding = pos_emb
        return st

    icases = []
    for num_side, patch, H_, W_, masks in (
            (5, 14, 70, 70, [np.ones((5, 5), bool)]),
            (7, 14, 70, 56, [np.ones((5, 4), bool), np.pad(np.ones((3, 2), bool), ((0, 2), (0, 2)))]),
            (70, 14, 98, 98, [np.pad(np.ones((7, 4), bool), ((0, 0), (0, 3))), np.ones((7, 7), bool)])):
        st = make_embed_stub(num_side, patch)
        x = np.zeros((len(masks), H_, W_, 3), np.float32)
        ns["VisionEmbeddings"].__call__(st, x, mask=[m for m in masks])
        icases.append({"num_patches_per_side": num_side, "patch_size": patch,
                       "patch_mask": [m.astype(int).tolist() for m in masks],
                       "position_ids": seen["ids"].tolist()})
    golden["idefics2_position_ids"] = icases

    ns["InputEmbeddingsFeatures"] = lambda inputs_embeds=None, **k: types.SimpleNamespace(inputs_embeds=inputs_embeds)
    load(ns, "models/idefics2/idefics2.py", "masked_scatter")
    gie, w = load(ns, "models/idefics2/idefics2.py", "get_input_embeddings", "Model")
    prep, w2 = load(ns, "models/idefics2/idefics2.py", "_prepare_inputs_for_multimodal", "Model")
    provenance["idefics2.get_input_embeddings"] = w
    provenance["idefics2._prepare_inputs_for_multimodal"] = w2
    cap = {}

    def vision_stub(x, patch_attention_mask=None, output_hidden_states=None):
        cap["n_images"] = int(x.shape[0])
        cap["patch_mask"] = np.asarray(patch_attention_mask).astype(int).tolist()
        cap["pixel_sum"] = [float(v) for v in np.asarray(x).reshape(x.shape[0], -1).sum(1)]
        return np.zeros((x.shape[0], 3, 2), np.float32), None, None

    mstub = types.SimpleNamespace(
        config=types.SimpleNamespace(vision_config=types.SimpleNamespace(patch_size=14), image_token_index=100),
        language_model=types.SimpleNamespace(embed_tokens=lambda ids: np.zeros(ids.shape + (2,), np.float32)),
        vision_model=vision_stub, connector=lambda f: f)
    mstub._prepare_inputs_for_multimodal = lambda feats, emb, ids: emb
    pvv = rng.standard_normal((1, 3, 3, 42, 56)).astype(np.float32)
    pvv[0, 1] = 0.0  # a padding image
    pam = np.zeros((1, 3, 42, 56), bool)
    pam[0, 0] = True
    pam[0, 2, :30, :20] = True
    gie(mstub, mx.array(np.zeros((1, 4), np.int64)), mx.array(pvv), pixel_attention_mask=mx.array(pam))
    golden["idefics2_get_input_embeddings"] = {"pixel_values_shape": list(pvv.shape), "zero_image": 1,
                                               "pixel_attention_valid": [[42, 56], [0, 0], [30, 20]],
                                               "n_images_kept": cap["n_images"], "patch_mask": cap["patch_mask"],
                                               "pixel_sum": cap["pixel_sum"],
                                               "pixel_sum_expected": [float(pvv[0, i].sum()) for i in (0, 2)]}
    pcases = []
    for ids_, nfeat, tag in (([[5, 100, 100, 6, 100]], 3, "three image rows"),
                             ([[5, 100, 6]], 2, "count mismatch -> ValueError")):
        ids_ = np.asarray(ids_)
        Hd = 3
        feats = (1000 + np.arange(nfeat * Hd, dtype=np.float32)).reshape(1, nfeat, Hd)
        emb = -(np.arange(ids_.size * Hd, dtype=np.float32) + 1).reshape(1, ids_.shape[1], Hd)
        try:
            out = tolist(prep(mstub, mx.array(feats), mx.array(emb), mx.array(ids_)))
            err = None
        except ValueError as e:
            out, err = None, str(e)
        pcases.append({"tag": tag, "input_ids": ids_.tolist(), "n_feats": nfeat, "hidden": Hd,
                       "output": out, "error": err})
    golden["idefics2_merge"] = pcases

    # ---------------- load path: key renames + conv layout rule (qwen2_vl.py:179-190, vision.py:9-25,292-310)
    msan, w = load(ns, "models/qwen2_vl/qwen2_vl.py", "sanitize", "Model")
    provenance["qwen2_vl.Model.sanitize"] = w
    hf_keys = ["visual.patch_embed.proj.weight", "visual.blocks.0.attn.qkv.weight", "visual.merger.mlp.0.bias",
               "model.embed_tokens.weight", "model.layers.3.self_attn.q_proj.bias", "model.norm.weight",
               "lm_head.weight", "vision_tower.blocks.1.norm1.weight",
               "language_model.model.layers.0.mlp.up_proj.weight", "language_model.lm_head.weight"]
    golden["sanitize_keys"] = {k: list(msan(None, {k: 0}).keys())[0] for k in hf_keys}
    load(ns, "models/qwen2_vl/vision.py", "check_array_shape")
    vsan, w = load(ns, "models/qwen2_vl/vision.py", "sanitize", "VisionModel")
    provenance["qwen2_vl.VisionModel.sanitize"] = w
    hfw = np.arange(8 * 3 * 2 * 4 * 4, dtype=np.float32).reshape(8, 3, 2, 4, 4)   # HF [O, C, T, H, W]
    vout = vsan(None, {"vision_tower.patch_embed.proj.weight": hfw, "vision_tower.blocks.0.attn.position_ids": 1,
                       "vision_tower.blocks.0.attn.qkv.weight": np.zeros((2, 2), np.float32)})
    mlxw = vout["vision_tower.patch_embed.proj.weight"]
    again = vsan(None, {"vision_tower.patch_embed.proj.weight": mlxw})["vision_tower.patch_embed.proj.weight"]
    golden["vision_sanitize"] = {"hf_shape": list(hfw.shape), "kept_keys": sorted(vout.keys()),
                                 "mlx_layout_shape": list(mlxw.shape), "mlx_layout": tolist(mlxw),
                                 "idempotent": bool(np.array_equal(again, mlxw))}

    # ---------------- image processor (qwen3_vl/processing_qwen3_vl.py:164-354): PIL bicubic resize,
    # rescale / normalise, temporal duplication, merge-group-major patch rows
    import hashlib
    ns["math"] = __import__("math")
    ns["Tuple"], ns["List"] = tuple, list
    ns["Image"] = __import__("PIL.Image", fromlist=["Image"])
    load(ns, "models/qwen3_vl/processing_qwen3_vl.py", "_smart_resize_image")
    load(ns, "models/qwen3_vl/processing_qwen3_vl.py", "_resize_video_frames")
    ns["ImageProcessingMixin"] = object
    code, w = extract("models/qwen3_vl/processing_qwen3_vl.py", "Qwen3VLImageProcessor")
    exec(compile(code, "<ref Qwen3VLImageProcessor>", "exec"), ns)
    provenance["Qwen3VLImageProcessor._process_one"] = w
    IP = ns["Qwen3VLImageProcessor"]
    ip = IP(patch_size=14, temporal_patch_size=2, merge_size=2,
            image_mean=[0.48145466, 0.4578275, 0.40821073], image_std=[0.26862954, 0.26130258, 0.27577711])
    pcases = []
    for hw_ in ((336, 336), (56, 84), (200, 310), (28, 1000), (30, 45)):
        img = np.random.default_rng(hw_[0] * 1000 + hw_[1]).integers(0, 256, size=(3, hw_[0], hw_[1]), dtype=np.uint8)
        pvv_, grid_ = ip._process_one(img)
        pvv_ = np.ascontiguousarray(pvv_.astype(np.float32))
        pcases.append({"hw": list(hw_), "seed": hw_[0] * 1000 + hw_[1], "grid": [int(x) for x in grid_],
                       "shape": list(pvv_.shape), "sha256": hashlib.sha256(pvv_.tobytes()).hexdigest(),
                       "first": tolist(pvv_[0, :4]), "last": tolist(pvv_[-1, -4:])})
    golden["image_processor"] = pcases

    # ---------------- streaming detokenizer (tokenizer_utils.py:14-118)
    ns["REPLACEMENT_CHAR"] = "\ufffd"
    code, w = extract("tokenizer_utils.py", "StreamingDetokenizer")
    exec(compile(code, "<ref StreamingDetokenizer>", "exec"), ns)
    code, w = extract("tokenizer_utils.py", "NaiveStreamingDetokenizer")
    exec(compile(code, "<ref NaiveStreamingDetokenizer>", "exec"), ns)
    provenance["NaiveStreamingDetokenizer"] = w
    pieces = {0: b"", 1: b"Hel", 2: b"lo", 3: b" w", 4: "\u00e9".encode()[:1], 5: "\u00e9".encode()[1:], 6: b"\n",
              7: b"next", 8: "\U0001f600".encode()[:2], 9: "\U0001f600".encode()[2:], 10: b"!", 11: b"<eos>"}

    class FakeTok:
        def decode(self, ids):
            return b"".join(pieces[i] for i in ids).decode("utf-8", errors="replace")
    det = ns["NaiveStreamingDetokenizer"](FakeTok())
    seq = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 6, 1]
    dtrace = []
    for tk in seq:
        det.add_token(tk, skip_special_token_ids=[11])
        dtrace.append({"token": tk, "segment": det.last_segment, "settled_tokens": list(det.tokens)})
    det.finalize()
    golden["detokenizer_trace"] = {"pieces_hex": {str(k): v.hex() for k, v in pieces.items()}, "skip": [11],
                                   "steps": dtrace, "final_segment": det.last_segment, "final_text": det.text,
                                   "final_tokens": list(det.tokens)}

    # ---------------- logits processors (sample_utils.py:390-475): repetition / presence / frequency
    class _At:
        def __init__(self, arr):
            self.arr = arr

        def __getitem__(self, idx):
            arr = self.arr

            class _Op:
                def subtract(self_, v):   # mlx `.at[idx].subtract(v)`: duplicates accumulate
                    out = np.array(arr, copy=True)
                    np.subtract.at(out, idx, v)
                    return out.view(AtArray)
            return _Op()

    class AtArray(np.ndarray):
        @property
        def at(self):
            return _At(self)

    pen = {}
    plog = (rng.standard_normal((1, 16)).astype(np.float32) * 3.0)
    hist = [3, 7, 7, 1, 12, 3, 3, 9, 15, 0, 7]
    pen["logits"], pen["tokens"] = tolist(plog), hist
    for nm, arg, ctx in (("make_repetition_penalty", 1.3, 20), ("make_repetition_penalty", 1.3, 4),
                         ("make_presence_penalty", 0.7, 20), ("make_presence_penalty", 0.7, 3),
                         ("make_frequency_penalty", 0.4, 20), ("make_frequency_penalty", 0.4, 6)):
        mk, w = load(ns, "sample_utils.py", nm)
        provenance[nm] = w
        out = mk(arg, ctx)(list(hist), np.array(plog, copy=True).view(AtArray))
        pen[f"{nm}({arg},{ctx})"] = tolist(np.asarray(out))
        pen[f"{nm}({arg},{ctx}) empty history"] = tolist(np.asarray(mk(arg, ctx)([], np.array(plog, copy=True).view(AtArray))))
    golden["logits_processors"] = pen

    # ---------------- config schema (models/qwen2_vl/config.py): field names, order, literal defaults
    csrc = open(os.path.join(REF, "models/qwen2_vl/config.py")).read()
    schema = {}
    for node in ast.parse(csrc).body:
        if isinstance(node, ast.ClassDef):
            rows = []
            for st in node.body:
                if isinstance(st, ast.AnnAssign):
                    rows.append([st.target.id, None if st.value is None else ast.literal_eval(st.value)])
            schema[node.name] = rows
    golden["qwen2_vl_config_schema"] = schema
    provenance["qwen2_vl config schema"] = "models/qwen2_vl/config.py (AnnAssign nodes)"

    # ---------------- sampler masks (sample_utils.py:149-345), fp32 on seeded logprobs
    ns["math"] = __import__("math")
    samp = {}
    lg = rng.standard_normal((2, 12)).astype(np.float32) * 2.0
    lp_ = lg - np.log(np.sum(np.exp(lg), axis=-1, keepdims=True))
    samp["logits"], samp["logprobs"] = tolist(lg), tolist(lp_)
    f, w = load(ns, "sample_utils.py", "_apply_top_k"); provenance["_apply_top_k"] = w
    samp["top_k_3"] = tolist(f(mx.array(lp_), 3))
    f, w = load(ns, "sample_utils.py", "apply_top_p"); provenance["apply_top_p"] = w
    samp["top_p_0.7"] = tolist(f(mx.array(lp_), 0.7))
    f, w = load(ns, "sample_utils.py", "_apply_min_p"); provenance["_apply_min_p"] = w
    samp["min_p_0.2"] = tolist(f(mx.array(lp_), 0.2, 1))
    samp["min_p_0.6_keep3"] = tolist(f(mx.array(lp_), 0.6, 3))
    f, w = load(ns, "sample_utils.py", "_top_n_sigma"); provenance["_top_n_sigma"] = w
    samp["top_n_sigma_1.0"] = tolist(f(mx.array(lg), 1.0))
    f, w = load(ns, "sample_utils.py", "apply_p_less"); provenance["apply_p_less"] = w
    samp["p_less_t0.8"] = tolist(f(mx.array(lg), 0.8))
    f, w = load(ns, "sample_utils.py", "_typical_p"); provenance["_typical_p"] = w
    samp["typical_p_0.6"] = tolist(f(mx.array(lp_), 0.6))
    golden["sampler_masks"] = json.loads(json.dumps(samp).replace("-Infinity", '"-inf"'))

    with open(OUT, "w") as f:
        json.dump(golden, f)
    print(f"wrote {OUT} ({os.path.getsize(OUT)} bytes); functions: {sorted(provenance)}")


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference tree not present: golden vectors can only be regenerated in the build container")
    main()
