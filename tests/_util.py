"""Shared helpers for the parity tests (tests may import oracle/; the product may not)."""
import numpy as np
import torch


def to_dev(x, dtype=torch.bfloat16):
    return torch.as_tensor(x).to(device="cuda", dtype=dtype).contiguous()


def cmp_bf16(got: torch.Tensor, want: torch.Tensor, name="", rel_l2=1e-3, max_mismatch=0.02,
             ulps=2):
    """got: bf16 device tensor; want: fp32 CPU tensor of bf16-representable values.
    Bar: relative L2 error <= rel_l2 (the north star's 1e-3), at most `max_mismatch`
    of the elements differ at all, and no element differs by more than `ulps` bf16 ulps
    of max(|want|, tiny)."""
    g = got.detach().float().cpu().reshape(-1)
    w = want.detach().float().cpu().reshape(-1)
    assert g.shape == w.shape, (name, g.shape, w.shape)
    assert torch.isfinite(g).all(), f"{name}: non-finite output"
    diff = (g - w).abs()
    denom = w.norm().item() + 1e-30
    rl2 = diff.norm().item() / denom
    mism = (diff > 0).float().mean().item()
    ulp = torch.maximum(w.abs(), torch.full_like(w, 1e-30)) * 2.0 ** -7
    worst = (diff / ulp).max().item()
    msg = f"{name}: rel_l2={rl2:.3e} mismatch={mism:.4f} worst_ulps={worst:.2f}"
    print(msg)
    assert rl2 <= rel_l2, msg
    assert mism <= max_mismatch, msg
    return rl2, mism, worst


def rl2(a, b):
    a = a.detach().float().cpu().reshape(-1)
    b = b.detach().float().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def cmp_noise(got, want_bf16, want_f32, name="", slack=1.5):
    """End-to-end comparison for a DEEP bf16 pipeline.

    A chain of bf16-rounded ops is chaotic: one element that rounds the other way
    (fp32 summation order, tensor-core accumulation) perturbs every downstream dot
    product and flips ~1/sqrt(K) of the next layer's roundings, so after a few
    layers two faithful implementations differ on ~half of the elements by one
    ulp.  The per-op bar (<= 1e-3, test_kernels_gpu.py) is therefore checked on
    identical inputs; end to end we require that the CUDA path is
      (a) no further from the oracle's bf16 result than `slack` (1.5 ~ sqrt(2), two
          independent noise realisations, plus margin) times the distance of the
          oracle's bf16 result from the exact (fp32, un-rounded) evaluation, and
      (b) as close to the exact evaluation as the oracle's bf16 result is.
    """
    d_ref = rl2(want_bf16, want_f32)   # the reference's own rounding noise
    d_pair = rl2(got, want_bf16)
    d_ours = rl2(got, want_f32)
    g = got.detach().float().cpu().reshape(-1)
    assert torch.isfinite(g).all(), f"{name}: non-finite output"
    msg = (f"{name}: |cuda-oracle_bf16|={d_pair:.3e} |oracle_bf16-exact|={d_ref:.3e} "
           f"|cuda-exact|={d_ours:.3e}")
    print(msg)
    assert d_pair <= max(1e-3, slack * d_ref), msg
    assert d_ours <= 1.25 * d_ref + 1e-3, msg
    return d_pair, d_ref, d_ours
