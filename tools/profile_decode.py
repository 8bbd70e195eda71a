"""Profiling target: C2 request, prefill once + a few decode steps as plain launches
(no graph) so that ncu lists every kernel.  usage: profile_decode.py [n_steps] [graph]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_vlm_b200.models.cache import make_prompt_cache
from mlx_vlm_b200.utils import load_synthetic, prepare_inputs
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
use_graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
model, proc = load_synthetic("qwen2-vl-2b", seed=0, device="cuda:0", n_text_tokens=128)
eng = model.engine
eng.set_graph(use_graph)
img = np.random.default_rng(0).integers(0, 256, size=(336, 336, 3), dtype=np.uint8)
inp = prepare_inputs(proc, images=[img], prompts="x", device=eng.device, stream=eng.stream)
ids, pvd, grid = inp["input_ids"], inp["pixel_values"], inp["image_grid_thw"]
T = ids.shape[1]
for it in range(2):
    cache = make_prompt_cache(model.language_model)
    emb = model.get_input_embeddings(ids, pvd, image_grid_thw=grid)
    model.language_model(ids, inputs_embeds=emb.inputs_embeds, cache=cache, position_ids=emb.position_ids,
                         rope_deltas=emb.rope_deltas, logits_to_keep=1, reserve_tokens=T + 600)
    # put ~400 tokens of context in the cache first (fast), then the profiled steps
    model.language_model.fused_greedy_decode_n(n_steps, cache, reserve_tokens=T + 600)
    eng.stream.synchronize()
print("done", eng.launch_count)
