"""Profiling target for the lock-step batched decoder: Qwen2-VL-7B shapes (n layers), B rows at ctx ~ C,
a few eager steps (no graph) so that ncu lists every kernel.  usage: profile_batch.py [layers] [B] [ctx] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_vlm_b200.models.qwen2_vl import Model
from mlx_vlm_b200.models.qwen2_vl.config import qwen2_vl_7b_config

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C = int(sys.argv[3]) if len(sys.argv) > 3 else 500
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
cfg = qwen2_vl_7b_config()
cfg.text_config.num_hidden_layers = L
cfg.vision_config.depth = 1
model = Model(cfg, device="cuda:0").init_random(0)
lm, eng = model.language_model, model.engine
rows, caches = lm.make_batch_cache(B, 1024)
rows.lengths = [C + 7 * b for b in range(B)]
rows._touch()
toks = np.arange(B) + 5
for s in range(steps):
    out = lm.fused_greedy_decode(toks[:, None], cache=caches, rope_deltas=np.zeros((B, 1), dtype=np.int64))
    eng.stream.synchronize()
    toks = out.cpu().numpy()
print("done", eng.launch_count, eng.last_decode_ms())
