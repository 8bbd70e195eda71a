"""Dev check of k_mega_tc internals on a 1-layer model: split-K partial sums and act
against fp32 torch GEMVs of the same weights.  usage: tc_check2.py [tiny1|wide1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_engine_gpu import _build
from _util import rl2
from mlx_vlm_b200.models.cache import make_prompt_cache
from oracle import qwen2vl as O

kind = sys.argv[1] if len(sys.argv) > 1 else "tiny1"
c, W, model, req = _build(kind, 16, (56, 56))
ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
eng = model.engine
t = c.text
H, I = t.hidden_size, t.intermediate_size
hd = H // t.num_attention_heads
QKV = (t.num_attention_heads + 2 * t.num_key_value_heads) * hd
pvd = torch.from_numpy(pv).cuda()
T = ids.shape[1]
tok = 7
def bf(x): return x.to(torch.bfloat16).float()
P = "language_model.model.layers.0."
def w(n): return W[n].float().cuda()
acts = {}
for mode in (1, 2):
    eng.set_mega(mode)
    cache = make_prompt_cache(model.language_model)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    out = model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=cache,
                               position_ids=emb.position_ids, rope_deltas=emb.rope_deltas)
    eng.stream.synchronize()
    eng.set_next(tok, T, T)
    eng.decode(1, force_tokens=np.asarray([3], dtype=np.int32))
    eng.stream.synchronize()
    print("mode", mode, "device_error", eng.device_error())
    acts[mode] = eng.debug_buffer("act", torch.bfloat16).float().cpu()
    if mode == 1:
        h1 = eng.debug_buffer("h", torch.bfloat16).float().cpu()
    logits = eng.logits_view().float().cpu().clone()
    acts[("logits", mode)] = logits
    acts[("k", mode)] = cache[0].keys[0, :, T].float().cpu().clone()
    acts[("v", mode)] = cache[0].values[0, :, T].float().cpu().clone()
print("act mode2 vs mode1 rl2", rl2(acts[2], acts[1]), "norms", float(acts[1].norm()), float(acts[2].norm()))
print("logits mode2 vs mode1 rl2", rl2(acts[("logits", 2)], acts[("logits", 1)]))
print("new K row mode2 vs mode1", rl2(acts[("k", 2)], acts[("k", 1)]), float(acts[("k", 1)].norm()), float(acts[("k", 2)].norm()))
print("new V row mode2 vs mode1", rl2(acts[("v", 2)], acts[("v", 1)]), float(acts[("v", 1)].norm()), float(acts[("v", 2)].norm()))
