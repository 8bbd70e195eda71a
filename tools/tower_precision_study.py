"""CPU study for the round-2 LLaVA / Idefics2 vision tower (DESIGN.md round-2 plan, item 4).

The reference computes the CLIP tower in fp32 (float32 pixel_values x bf16-valued weights) and rounds
the projected features to bf16 only at the merge.  Which tensor-core formulation reproduces THOSE
bf16 features?  Compared here, on the real CLIP-L/14 widths (1024 / 16 heads / 4096, N layers) with
seeded random weights, against the fp32 tower (oracle/llava.py, Rounder("f32")):
  bf16      activations rounded to bf16 after every op (what the Qwen2-VL kernels do)
  split-2   every GEMM input x = hi + lo (two bf16 terms), hi*hi + hi*lo + lo*hi in fp32 — 3 MMAs
  split-3   three bf16 terms, all products above 2^-24                                   — 6 MMAs
  tf32      GEMM inputs rounded to 10 mantissa bits (tcgen05 kind::tf32), fp32 everywhere else
Reported: relative L2 of the projected features and the fraction of bf16-rounded feature elements
that differ from the reference's.   usage: tower_precision_study.py [layers] [image_size]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import llava as L
from oracle import mlx_semantics as S
from oracle.mlx_semantics import Rounder

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 24
image = int(sys.argv[2]) if len(sys.argv) > 2 else 168


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def tf32(x):
    # round-to-nearest-even to 10 explicit mantissa bits
    i = x.contiguous().view(torch.int32)
    r = ((i >> 13) & 1) + 0x0FFF
    return ((i + r) & ~0x1FFF).view(torch.float32)


def make_linear(mode):
    def lin(R, x, w, b=None):
        if mode == "tf32":
            y = tf32(x) @ tf32(w).T
        else:
            n = int(mode[-1])
            xs, rem = [], x
            for _ in range(n):
                h = bf(rem)
                xs.append(h)
                rem = rem - h
            # weights are bf16-valued already: one term
            y = sum(t @ w.T for t in xs)
        return y if b is None else y + b
    return lin


cfg = L.llava_15_7b()
cfg.vision.num_hidden_layers = layers
cfg.vision.image_size = image
cfg.text = L.LlamaCfg(hidden_size=4096, num_hidden_layers=0, intermediate_size=64, num_attention_heads=32,
                      num_key_value_heads=32, vocab_size=64)
torch.manual_seed(0)
W = L.init_weights(cfg, seed=0)
pv = torch.randn(1, image, image, 3)
ref = L.image_features(cfg, W, pv, Rounder("f32"))
ref_b = bf(ref)
print(f"CLIP-L/14 widths, {layers} layers, {image}x{image} image ({cfg.vision.num_patches} patches); "
      f"reference = fp32 tower, features rounded to bf16 at the merge")


def report(name, out):
    rel = float((out - ref).norm() / ref.norm())
    flips = float((bf(out) != ref_b).float().mean())
    print(f"  {name:8s} rel_l2 vs fp32 {rel:.3e}   bf16-rounded features that differ: {100 * flips:.3f} %")


report("bf16", L.image_features(cfg, W, pv, Rounder("bf16")))
orig = S.linear
for mode in ("split2", "split3", "tf32"):
    S.linear = make_linear(mode)
    try:
        report(mode, L.image_features(cfg, W, pv, Rounder("f32")))
    finally:
        S.linear = orig
