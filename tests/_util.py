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
