"""Dev check: k_mega_tc (modes 2/3) against k_mega (mode 1) and the oracle, teacher-forced.
usage: tc_check.py [kind] [n_dec]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_engine_gpu import _build
from _util import rl2
from mlx_vlm_b200.models.cache import make_prompt_cache
from oracle import qwen2vl as O

kind = sys.argv[1] if len(sys.argv) > 1 else "wide2"
n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c, W, model, req = _build(kind, 16, (56, 56))
ids, pv, grid = req["input_ids"], req["pixel_values"], req["image_grid_thw"]
eng = model.engine
ref = O.greedy_generate(c, W, ids, pv, grid, n_dec)
toks = ref["tokens"][0].tolist()
ex = O.greedy_generate(c, W, ids, pv, grid, n_dec, dtype="f32", force_tokens=toks[:])
pvd = torch.from_numpy(pv).cuda()
T = ids.shape[1]
res = {}
for mode in (1, 2, 3):
    eng.set_mega(mode)
    cache = make_prompt_cache(model.language_model)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    out = model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=cache,
                               position_ids=emb.position_ids, rope_deltas=emb.rope_deltas)
    eng.stream.synchronize()
    delta = int(ref["prefill"].rope_deltas[0, 0])
    eng.set_next(toks[0], T, T + delta)
    logs = []
    for n in range(1, n_dec):
        eng.decode(1, force_tokens=np.asarray([toks[n]], dtype=np.int32))
        eng.stream.synchronize()
        logs.append(eng.logits_view().float().cpu().clone())
    err = eng.device_error()
    kc = cache[0].keys[0, :, :T + n_dec - 1].float().cpu().clone()
    kl = cache[-1].keys[0, :, :T + n_dec - 1].float().cpu().clone()
    vl = cache[-1].values[0, :, :T + n_dec - 1].float().cpu().clone()
    res[mode] = (logs, kc, kl, vl)
    print(f"mode {mode}: device_error={err}")
    for n, lg in enumerate(logs, 1):
        o_bf, o_32 = ref["logits"][n][0].float(), ex["logits"][n][0].float()
        print(f"  step {n}: |m-oracle_bf16|={rl2(lg, o_bf):.3e} |oracle_bf16-exact|={rl2(o_bf, o_32):.3e} "
              f"|m-exact|={rl2(lg, o_32):.3e} nan={int(torch.isnan(lg).sum())} argmax={int(lg.argmax())} oracle={int(o_bf.argmax())}")
for mode in (2, 3):
    for n in range(len(res[1][0])):
        print(f"mode {mode} vs 1 step {n+1}: rl2={rl2(res[mode][0][n], res[1][0][n]):.3e}")
    print(f"mode {mode} K cache L0 vs mode1: {rl2(res[mode][1], res[1][1]):.3e}  last-layer K {rl2(res[mode][2], res[1][2]):.3e} V {rl2(res[mode][3], res[1][3]):.3e}")
