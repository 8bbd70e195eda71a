"""A/B of two builds of libb200vlm.so on the same box: python tools/wt_ab_probe.py libA.so libB.so
Times gemm_wt (graph replay of 24 launches) for a few shapes / modes / debug flags with identical buffers."""
import ctypes as C
import sys

import torch

dev = "cuda:0"
P, L, I, U = C.c_void_p, C.c_long, C.c_int, C.c_uint


def load(path):
    lib = C.CDLL(path)
    lib.b200_gemm_wt.restype = I
    lib.b200_gemm_wt.argtypes = [P, L, P, P, P, L, P, L, P, I, I, I, I, I, I, C.POINTER(I), U, P]
    lib.b200_last_error.restype = C.c_char_p
    return lib


def bf(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed + sum(shape))
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)


def time_cfg(lib, T, Nn, K, mode, inter, cfg, flags, rep=24, bias=False, epi=0):
    X = [bf(T, K, seed=i) for i in range(2)]
    W = [bf(Nn, K, scale=0.03, seed=i) for i in range(6)]
    ncol = inter if mode == 2 else Nn
    out = torch.empty(T, ncol, device=dev, dtype=torch.bfloat16)
    part = torch.empty(max(cfg[3], 1), T, Nn, device=dev, dtype=torch.float32) if mode == 1 else None
    s = torch.cuda.Stream()
    b = bf(Nn, scale=0.5) if bias else None

    def call(i):
        rc = lib.b200_gemm_wt(X[i & 1].data_ptr(), K, W[i % 6].data_ptr(), b.data_ptr() if bias else None, None, 0, out.data_ptr(), ncol,
                              part.data_ptr() if part is not None else None, T, Nn, K, epi, mode, inter, (I * 4)(*cfg), flags,
                              s.cuda_stream)
        assert rc == 0, lib.b200_last_error()
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


libs = [(p, load(p)) for p in sys.argv[1:]]
cases = [("vit.qkv", 576, 3840, 1280, 0, 0, (144, 2, 3, 1)), ("vit.fc1", 576, 5120, 1280, 0, 0, (192, 2, 2, 1)),
         ("lm.gateup", 272, 17920, 1536, 2, 8960, (144, 1, 3, 1)), ("vit.proj", 576, 1280, 1280, 1, 0, (96, 2, 3, 2)),
         ("lm.qkv", 272, 2048, 1536, 0, 0, (96, 2, 3, 1))]
for rnd in range(2):
    for name, T, Nn, K, mode, inter, cfg in cases:
        for path, lib in libs:
            r = [time_cfg(lib, T, Nn, K, mode, inter, cfg, fl) for fl in (0, 1, 2, 4)]
            extra = ""
            if mode == 0:
                extra = f" | bias {time_cfg(lib, T, Nn, K, mode, inter, cfg, 0, bias=True):6.2f} | bias+gelu {time_cfg(lib, T, Nn, K, mode, inter, cfg, 0, bias=True, epi=1):6.2f}"
            print(f"{name:10s} {cfg} {path[-28:]:28s}: full {r[0]:6.2f} | loads-only {r[1]:6.2f} | mma-only {r[2]:6.2f} | no-store {r[3]:6.2f}{extra} us",
                  flush=True)
