"""gemm_wt (weight-major tcgen05 GEMM) on the GPU: correctness against a torch fp32 evaluation of
the same rounding points, then a timing sweep over tile / pipeline configurations at the prefill
shapes (CUDA-graph replay of 24 launches over rotating weight buffers, CUDA events).
usage: python tools/wt_sweep.py [check|sweep|all] [quick]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mlx_vlm_b200 import _native as N

lib = N.lib()
dev = "cuda:0"
what = sys.argv[1] if len(sys.argv) > 1 else "all"
quick = len(sys.argv) > 2
BF16, PARTIAL, SWIGLU = 0, 1, 2


def bf(*shape, scale=1.0, seed=None):
    g = torch.Generator(device=dev)
    g.manual_seed(seed if seed is not None else 1234 + sum(shape))
    return (torch.randn(*shape, device=dev, dtype=torch.float32, generator=g) * scale).to(torch.bfloat16)


def r(x):
    return x.to(torch.bfloat16).float()


def cfg_arr(cfg):
    return (C.c_int * 4)(*cfg) if cfg else None


def call(X, W, bias, resid, Cout, partial, T, Nn, K, epi, mode, inter, cfg, flags=0, stream=0):
    N.check(lib.b200_gemm_wt(X.data_ptr(), X.stride(0), W.data_ptr(), N.ptr(bias), N.ptr(resid),
                             resid.stride(0) if resid is not None else 0, N.ptr(Cout),
                             Cout.stride(0) if Cout is not None else 0, N.ptr(partial), T, Nn, K, epi, mode,
                             inter, cfg_arr(cfg), flags, stream), "gemm_wt")


def gelu_fast(x):
    return r(x * r(torch.sigmoid(r(1.703125 * x))))


def swiglu(g, u):
    return r(r(g * r(torch.sigmoid(g))) * u)


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def check():
    ok = True
    cases = [  # T, N, K, epi, bias, resid, cfg
        (576, 3840, 1280, 0, True, False, None), (576, 5120, 1280, 1, True, False, (192, 2, 2, 1)),
        (576, 1280, 5120, 0, True, True, (192, 1, 4, 1)), (272, 2048, 1536, 0, True, False, (144, 2, 3, 1)),
        (272, 1536, 8960, 0, False, True, None), (1, 256, 64, 0, True, False, (16, 1, 3, 1)),
        (130, 200, 72, 1, True, True, (144, 1, 2, 1)), (8, 1536, 1536, 0, True, False, (16, 2, 6, 1)),
        (300, 1000, 1176, 0, False, False, (80, 2, 4, 1)), (4616, 1024, 1024, 0, True, True, None)]
    for T, Nn, K, epi, hb, hr, cfg in cases:
        X, W = bf(T, K), bf(Nn, K, scale=0.05)
        b = bf(Nn, scale=0.5) if hb else None
        ldn = (Nn + 7) // 8 * 8
        R_ = bf(T, ldn) if hr else None
        Cc = torch.zeros(T, ldn, device=dev, dtype=torch.bfloat16)
        call(X, W, b, R_, Cc, None, T, Nn, K, epi, BF16, 0, cfg)
        torch.cuda.synchronize()
        want = X.float() @ W.float().t()
        if hb:
            want = want + b.float()
        want = r(want)
        if epi == 1:
            want = gelu_fast(want)
        if hr:
            want = r(R_[:, :Nn].float() + want)
        e = rel(Cc[:, :Nn], want)
        mism = float((Cc[:, :Nn].float() != want).float().mean())
        print(f"check BF16 T={T} N={Nn} K={K} epi={epi} cfg={cfg}: rel={e:.2e} mismatch={mism:.4f}", flush=True)
        ok &= e < 2e-3
    # split-K partial + finish_rows (RMS and LayerNorm) vs torch
    for T, Nn, K, cfg, kind in [(272, 1536, 8960, (144, 2, 3, 6), 1), (576, 1280, 5120, (192, 2, 2, 4), 2),
                                (8, 1536, 8960, (16, 2, 6, 10), 1), (272, 1536, 1536, (144, 2, 3, 3), 0)]:
        X, W = bf(T, K), bf(Nn, K, scale=0.05)
        b, h = bf(Nn, scale=0.5), bf(T, Nn)
        nw, nb = bf(Nn, scale=0.3) + 1.0, bf(Nn, scale=0.1)
        P = torch.full((cfg[3], T, Nn), 7.0, device=dev, dtype=torch.float32)
        call(X, W, None, None, None, P, T, Nn, K, 0, PARTIAL, 0, cfg)
        hout, xn = torch.zeros_like(h), torch.zeros_like(h)
        N.check(lib.b200_finish_rows(P.data_ptr(), cfg[3], b.data_ptr(), h.data_ptr(), Nn, hout.data_ptr(), Nn, kind,
                                     nw.data_ptr(), nb.data_ptr() if kind == 2 else 0, 1e-6, xn.data_ptr(), Nn, T, Nn, 0),
                "finish")
        torch.cuda.synchronize()
        lin = r(X.float() @ W.float().t() + b.float())
        hw = r(h.float() + lin)
        e1 = rel(hout, hw)
        hh = hout.float()
        if kind == 1:
            rs = torch.rsqrt((hh * hh).mean(-1, keepdim=True) + 1e-6)
            xw = r(r(hh * rs) * nw.float())
        elif kind == 2:
            mu = hh.mean(-1, keepdim=True)
            var = ((hh - mu) ** 2).mean(-1, keepdim=True)
            xw = r(r(r((hh - mu) * torch.rsqrt(var + 1e-6)) * nw.float()) + nb.float())
        else:
            xw = xn.float()
        e2 = rel(xn, xw)
        print(f"check PARTIAL T={T} N={Nn} K={K} cfg={cfg} norm={kind}: h rel={e1:.2e} xn rel={e2:.2e}", flush=True)
        ok &= e1 < 2e-3 and e2 < 2e-3
    # SwiGLU
    for T, I, K, cfg in [(272, 8960, 1536, None), (8, 8960, 1536, (16, 2, 6, 1)), (100, 200, 64, (112, 1, 2, 1))]:
        X, W = bf(T, K), bf(2 * I, K, scale=0.05)
        act = torch.zeros(T, I, device=dev, dtype=torch.bfloat16)
        call(X, W, None, None, act, None, T, 2 * I, K, 0, SWIGLU, I, cfg)
        torch.cuda.synchronize()
        gu = r(X.float() @ W.float().t())
        want = swiglu(gu[:, :I], gu[:, I:])
        e = rel(act, want)
        print(f"check SWIGLU T={T} I={I} K={K} cfg={cfg}: rel={e:.2e}", flush=True)
        ok &= e < 2e-3
    print("CHECK", "OK" if ok else "FAILED", flush=True)
    return ok


NBUF, REP = 6, 24


def time_cfg(T, Nn, K, mode, inter, cfg, flags=0):
    X = [bf(T, K) for _ in range(2)]
    W = [bf(Nn, K, scale=0.03, seed=i) for i in range(NBUF)]
    ncol = inter if mode == SWIGLU else Nn
    Cc = torch.empty(T, ncol, device=dev, dtype=torch.bfloat16)
    P = torch.empty(max(cfg[3], 1), T, Nn, device=dev, dtype=torch.float32) if mode == PARTIAL else None
    s = torch.cuda.Stream()
    try:
        with torch.cuda.stream(s):
            for i in range(3):
                call(X[i & 1], W[i % NBUF], None, None, Cc, P, T, Nn, K, 0, mode, inter, cfg, flags, s.cuda_stream)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for i in range(REP):
                    call(X[i & 1], W[i % NBUF], None, None, Cc, P, T, Nn, K, 0, mode, inter, cfg, flags, s.cuda_stream)
            g.replay()
            s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(4):
                g.replay()
            e1.record(s)
            s.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (4 * REP)
    except Exception as e:  # noqa
        print("   cfg", cfg, "failed:", str(e)[:120])
        return float("nan")


def sweep():
    shapes = [("vit.qkv", 576, 3840, 1280, BF16, 0), ("vit.proj", 576, 1280, 1280, PARTIAL, 0),
              ("vit.fc1", 576, 5120, 1280, BF16, 0), ("vit.fc2", 576, 1280, 5120, PARTIAL, 0),
              ("lm.qkv", 272, 2048, 1536, BF16, 0), ("lm.o", 272, 1536, 1536, PARTIAL, 0),
              ("lm.gateup", 272, 17920, 1536, SWIGLU, 8960), ("lm.down", 272, 1536, 8960, PARTIAL, 0)]
    if not quick:
        shapes += [("7b.gateup", 272, 37888, 3584, SWIGLU, 18944), ("7b.down", 272, 3584, 18944, PARTIAL, 0),
                   ("dec8.gateup", 8, 17920, 1536, SWIGLU, 8960), ("dec8.down", 8, 1536, 8960, PARTIAL, 0),
                   ("dec8.qkv", 8, 2048, 1536, PARTIAL, 0), ("dec8.7b.gateup", 8, 37888, 3584, SWIGLU, 18944),
                   ("clip.fc1x8", 4616, 4096, 1024, BF16, 0)]
    for name, T, Nn, K, mode, inter in shapes:
        auto = (C.c_int * 4)()
        lib.b200_gemm_wt_auto_config(T, Nn, K, mode, inter, auto)
        auto = tuple(auto)
        tns = sorted({auto[0]} | ({144, 272 // 2 // 16 * 16, 96} if T == 272 else {192, 144, 96} if T == 576 else {16} if T <= 16 else {256, 128}))
        tns = [t for t in tns if 16 <= t <= 256]
        rbs = (inter // 64) if mode == SWIGLU else (Nn + 127) // 128
        kb = (K + 63) // 64
        cands = [auto]
        for tn in tns:
            tt = (T + tn - 1) // tn
            for ks in (1, 2, 4):
                stage = ks * (16384 + tn * 128)
                for budget in (108 * 1024, 208 * 1024):
                    st = min(budget // stage, 12)
                    if st < 2:
                        continue
                    splits = [1]
                    if mode == PARTIAL:
                        base = rbs * tt
                        splits = sorted({max(1, min(kb // 4, x // base)) for x in (148, 296, 444)} | {1})
                    for sp in splits:
                        per = (kb + sp - 1) // sp
                        sp2 = (kb + per - 1) // per
                        cands.append((tn, ks, st, sp2))
        cands = list(dict.fromkeys(cands))
        res = []
        for cfg in cands:
            us = time_cfg(T, Nn, K, mode, inter, cfg)
            res.append((us, cfg))
        res.sort(key=lambda x: (x[0] != x[0], x[0]))
        fl = 2.0 * T * Nn * K
        best = res[0]
        au = [u for u, c in res if c == auto][0]
        print(f"{name:14s} T={T} N={Nn} K={K} mode={mode}: best {best[0]:7.2f} us {best[1]} ({fl / best[0] / 1e6:6.1f} TF, "
              f"W {Nn * K * 2 / best[0] / 1e3:5.0f} GB/s) | auto {auto} {au:7.2f} us", flush=True)
        print("     top: " + "  ".join(f"{u:.1f}:{c}" for u, c in res[:6]), flush=True)
        # bottleneck probes on the best config: loads only / MMA only / no epilogue stores
        b = best[1]
        print(f"     probes on {b}: loads-only {time_cfg(T, Nn, K, mode, inter, b, 1):.2f} us, "
              f"mma-only {time_cfg(T, Nn, K, mode, inter, b, 2):.2f} us, no-store {time_cfg(T, Nn, K, mode, inter, b, 4):.2f} us",
              flush=True)


if __name__ == "__main__":
    good = True
    if what in ("check", "all"):
        good = check()
    if what in ("sweep", "all") and good:
        sweep()
