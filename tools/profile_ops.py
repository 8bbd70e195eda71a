"""Profiling target for single hot ops with explicit configurations (ncu --set full -k regex:... -s 2 -c 1):
   python tools/profile_ops.py vit_fc1 | vit_qkv | lm_gateup | dec8_gateup | clip_fc1x8 | fa_vit | fa_lm
Each op runs 4 times (2 warm-up launches to skip)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mlx_vlm_b200 import _native as N

lib = N.lib()
dev = "cuda:0"
BF16, PARTIAL, SWIGLU = 0, 1, 2
op = sys.argv[1] if len(sys.argv) > 1 else "vit_fc1"


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


GEMMS = {   # T, N, K, epilogue, mode, inter, cfg (TN, KS, stages, split)  -- the measured winners (profiles/r2_gemm_wt_sweep.txt)
    "vit_qkv": (576, 3840, 1280, 0, BF16, 0, (144, 2, 3, 1)),
    "vit_fc1": (576, 5120, 1280, 1, BF16, 0, (192, 2, 2, 1)),
    "lm_gateup": (272, 17920, 1536, 0, SWIGLU, 8960, (144, 1, 3, 1)),
    "dec8_gateup": (8, 37888, 3584, 0, SWIGLU, 18944, (16, 2, 3, 1)),
    "clip_fc1x8": (4616, 4096, 1024, 1, BF16, 0, (256, 1, 2, 1)),
}
if op in GEMMS:
    T, Nn, K, epi, mode, inter, cfg = GEMMS[op]
    X, W, b = bf(T, K), bf(Nn, K, scale=0.05), bf(Nn, scale=0.5)
    out_cols = inter if mode == SWIGLU else Nn
    out = torch.zeros(T, out_cols, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        N.check(lib.b200_gemm_wt(X.data_ptr(), X.stride(0), W.data_ptr(), b.data_ptr() if mode == BF16 else 0, 0, 0,
                                 out.data_ptr(), out.stride(0), 0, T, Nn, K, epi, mode, inter, (C.c_int * 4)(*cfg), 0, 0),
                "gemm_wt")
    torch.cuda.synchronize()
    print(op, "done", float(out.float().abs().mean()))
elif op in ("fa_vit", "fa_lm"):
    nh, nkv, hd, L, causal = (16, 16, 80, 576, 0) if op == "fa_vit" else (12, 2, 128, 272, 1)
    q, k = bf(L, nh * hd), bf(L, nkv * hd)
    Lp = (L + 7) // 8 * 8
    vt = bf(nkv * hd, Lp)
    o = torch.zeros(L, nh * hd, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        N.check(lib.b200_attention_fa(q.data_ptr(), nh * hd, hd, k.data_ptr(), nkv * hd, hd, vt.data_ptr(), hd * Lp, Lp,
                                      o.data_ptr(), nh * hd, nh, nkv, hd, L, L, causal, 0), "attention_fa")
    torch.cuda.synchronize()
    print(op, "done", float(o.float().abs().mean()))
else:
    raise SystemExit(f"unknown op {op}")
