"""Does the bias / GELU epilogue of gemm_wt cost what the in-engine measurement suggests?  Times the same GEMM
(graph replay of 24 launches, rotating weights) without bias, with bias, with bias + fast GELU."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mlx_vlm_b200 import _native as N

lib = N.lib()
dev = "cuda:0"


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


def time_cfg(T, Nn, K, cfg, bias, epi, rot=6, rep=24):
    X = [bf(T, K, seed=i) for i in range(2)]
    W = [bf(Nn, K, scale=0.03, seed=i) for i in range(rot)]
    b = bf(Nn, scale=0.5) if bias else None
    out = torch.empty(T, Nn, device=dev, dtype=torch.bfloat16)
    s = torch.cuda.Stream()

    def call(i):
        N.check(lib.b200_gemm_wt(X[i & 1].data_ptr(), K, W[i % rot].data_ptr(), N.ptr(b), 0, 0, out.data_ptr(), Nn, 0, T, Nn,
                                 K, epi, 0, 0, (C.c_int * 4)(*cfg), 0, s.cuda_stream), "gemm_wt")
    with torch.cuda.stream(s):
        for i in range(3):
            call(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(rep):
                call(i)
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(4):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (4 * rep)


for name, T, Nn, K, cfgs in (("vit.qkv", 576, 3840, 1280, [(144, 2, 3, 1), (192, 2, 2, 1)]),
                             ("vit.fc1", 576, 5120, 1280, [(192, 2, 2, 1), (96, 1, 3, 1)]),
                             ("lm.qkv", 272, 2048, 1536, [(96, 2, 3, 1)]),
                             ("vit.patch", 576, 1280, 1176, [(96, 2, 3, 1), (144, 2, 3, 1)])):
    for cfg in cfgs:
        for rot in (6, 1):
            r = [time_cfg(T, Nn, K, cfg, bias, epi, rot=rot) for bias, epi in ((False, 0), (True, 0), (True, 1))]
            print(f"{name:10s} cfg={cfg} rotating W={rot}: no-bias {r[0]:6.2f} us | bias {r[1]:6.2f} us | bias+gelu {r[2]:6.2f} us", flush=True)
