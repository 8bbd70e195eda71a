"""Profiling target for the fp32-accurate CLIP tower of LLaVA-1.5 (C3): CLIP-L/14-336 widths, n layers, B images,
a few calls of encode_image so that ncu lists every kernel.  usage: profile_llava.py [layers] [B] [calls]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mlx_vlm_b200.models.llava import Model
from mlx_vlm_b200.models.llava.config import ModelConfig, TextConfig, VisionConfig

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = ModelConfig(text_config=TextConfig(hidden_size=4096, num_hidden_layers=1, intermediate_size=1024, num_attention_heads=32,
                                         num_key_value_heads=32, vocab_size=32064),
                  vision_config=VisionConfig(num_hidden_layers=L), vision_feature_layer=-1)
model = Model(cfg, device="cuda:0").init_random(0)
eng = model.engine
pv = torch.randn(B, 3, 336, 336, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(calls):
    ev[0].record(eng.stream)
    f = model.encode_image(pv)
    ev[1].record(eng.stream)
    eng.stream.synchronize()
    print(f"call {i}: {ev[0].elapsed_time(ev[1]):.3f} ms for {L} blocks x {B} images", flush=True)
