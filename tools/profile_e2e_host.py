"""Where does the host time of generate(model, processor, prompt, image=<host array>) go?  cProfile of one warm call
(C2 shapes) + wall-clock split: preprocessing / embeddings+prefill enqueue / decode loop / tail."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_vlm_b200 import generate
from mlx_vlm_b200.utils import load_synthetic

model, proc = load_synthetic("qwen2-vl-2b", seed=0, device="cuda:0", n_text_tokens=128)
model.config.eos_token_id = []
img = np.random.default_rng(0).integers(0, 256, size=(336, 336, 3), dtype=np.uint8)
for _ in range(3):
    generate(model, proc, "Describe this image in detail.", image=[img], max_tokens=512)
torch.cuda.synchronize()
t = time.perf_counter()
r = generate(model, proc, "Describe this image in detail.", image=[img], max_tokens=512)
torch.cuda.synchronize()
print(f"wall {1e3 * (time.perf_counter() - t):.1f} ms, generation_tps {r.generation_tps:.1f}, prompt_tps {r.prompt_tps:.1f}")
pr = cProfile.Profile()
pr.enable()
generate(model, proc, "Describe this image in detail.", image=[img], max_tokens=512)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
print(s.getvalue()[:6000])
