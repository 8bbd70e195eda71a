"""attention_fa vs attention_tc timing at the C2 shapes (CUDA-graph replay, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlx_vlm_b200 import _native as N
lib = N.lib()
dev = "cuda:0"


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def timeit(fn, rep=24):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(s.cuda_stream)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(rep):
                fn(s.cuda_stream)
        g.replay(); s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(4):
            g.replay()
        e1.record(s); s.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (4 * rep)


for name, nh, nkv, hd, L, S, causal in [("vit", 16, 16, 80, 576, 576, 0), ("lm", 12, 2, 128, 272, 272, 1),
                                        ("lm7b", 28, 4, 128, 272, 272, 1), ("clip", 16, 16, 64, 577, 577, 0),
                                        ("lm2k", 12, 2, 128, 2048, 2048, 1)]:
    q = bf(L, nh * hd); k = bf(S, nkv * hd); v = bf(nkv, S, hd)
    s_ld = (S + 7) // 8 * 8
    vt = torch.zeros(nkv, hd, s_ld, device=dev, dtype=torch.bfloat16); vt[:, :, :S] = v.transpose(1, 2)
    kk = k.view(S, nkv, hd).transpose(0, 1).contiguous()
    o = torch.empty(L, nh * hd, device=dev, dtype=torch.bfloat16)
    t_fa = timeit(lambda st: N.check(lib.b200_attention_fa(q.data_ptr(), nh * hd, hd, k.data_ptr(), nkv * hd, hd, vt.data_ptr(),
                                                          hd * s_ld, s_ld, o.data_ptr(), nh * hd, nh, nkv, hd, L, S, causal, st), "fa"))
    t_tc = timeit(lambda st: N.check(lib.b200_attention(q.data_ptr(), nh * hd, hd, kk.data_ptr(), hd, S * hd, v.data_ptr(), hd, S * hd,
                                                        o.data_ptr(), nh * hd, nh, nkv, hd, L, S, causal, 1.0 / hd ** 0.5, st), "tc"))
    fl = 4.0 * L * S * hd * nh * (0.5 if causal else 1.0)
    print(f"attn {name:5s} h={nh} kv={nkv} hd={hd} L={L}: fa {t_fa:7.2f} us ({fl / t_fa / 1e6:6.1f} TF)   tc(r1) {t_tc:7.2f} us", flush=True)
