"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the `mlx` (third-party, NOT vendored under /root/reference;
pinned mlx==0.32.0 / mlx-cpu==0.30.4, reference uv.lock:1094-1095,1160-1161) op
semantics that the reference's generate path relies on when it runs on the MLX
*CPU* device.  Every function below states WHERE it rounds to the activation
dtype, because the reference model code is dtype-generic and the rounding points
are what defines "the reference's bf16 output".

PARITY STATUS: **unpinned at the mlx boundary** — `import mlx` fails in this
container (no wheel, no network), so these semantics are restated from the
published mlx sources (mlx/fast.cpp fallback graphs, python/mlx/nn/layers/*.py)
as remembered, not checked against a running mlx.  The integer parts of the
path (rope index, merge indexing, cache bookkeeping) ARE pinned against the
reference's own tests (tests/test_oracle_golden.py).

Conventions: tensors are torch float32 CPU tensors whose values are exactly
representable in the activation dtype ("act dtype": bf16, fp16 or f32).
`Rounder(dtype).r(x)` rounds an fp32 tensor to the act dtype and returns fp32.
Matmuls accumulate in fp32 (mlx CPU gemm for 16-bit types accumulates in float
and casts the result once).
"""
from __future__ import annotations

import math

import torch

_DT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


class Rounder:
    """Rounds fp32 tensors to the activation dtype (RNE) and back to fp32."""

    def __init__(self, dtype: str = "bf16"):
        assert dtype in _DT
        self.name = dtype
        self.dtype = _DT[dtype]

    def r(self, x: torch.Tensor) -> torch.Tensor:
        if self.dtype == torch.float32:
            return x
        return x.to(self.dtype).to(torch.float32)

    def scalar(self, v: float) -> float:
        """A python scalar that meets an act-dtype array is weak-typed: it is
        converted to the array dtype before the op (mlx type promotion)."""
        if self.dtype == torch.float32:
            return float(torch.tensor(v, dtype=torch.float32))
        return float(torch.tensor(v, dtype=self.dtype).to(torch.float32))


# ---------------------------------------------------------------------------
# nn.Linear / nn.Embedding.as_linear  (python/mlx/nn/layers/linear.py:
#   `mx.addmm(bias, x, W.T)` when bias else `x @ W.T`)  -> ONE rounding.
# ---------------------------------------------------------------------------
def linear(R: Rounder, x, w, b=None):
    y = x @ w.T
    if b is not None:
        y = y + b
    return R.r(y)


# ---------------------------------------------------------------------------
# mx.fast.rms_norm on the CPU device = the fallback graph in mlx/fast.cpp:
#   x32 = astype(x, f32); x32 = x32 * rsqrt(mean(x32^2) + eps)
#   y = astype(x32, out); y = weight * y          -> TWO roundings.
# ---------------------------------------------------------------------------
def rms_norm(R: Rounder, x, w, eps: float):
    n = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return R.r(R.r(n) * w)


# ---------------------------------------------------------------------------
# mx.fast.layer_norm CPU fallback: fp32 statistics, cast, then *weight, +bias in
# the act dtype                                             -> THREE roundings.
# ---------------------------------------------------------------------------
def layer_norm(R: Rounder, x, w, b, eps: float):
    mu = x.mean(-1, keepdim=True)
    xc = x - mu
    var = xc.pow(2).mean(-1, keepdim=True)
    y = R.r(xc * torch.rsqrt(var + eps))
    if w is not None:
        y = R.r(y * w)
    if b is not None:
        y = R.r(y + b)
    return y


# ---------------------------------------------------------------------------
# Activations (python/mlx/nn/layers/activations.py).  They are `mx.compile`d but a
# compiled mlx kernel keeps every intermediate in its own dtype, so each
# primitive rounds.
# ---------------------------------------------------------------------------
def silu(R: Rounder, x):
    """nn.silu(x) = x * mx.sigmoid(x)"""
    return R.r(x * R.r(torch.sigmoid(x)))


def swiglu(R: Rounder, gate, up):
    """reference models/activations.py:8-10: nn.silu(gate) * x"""
    return R.r(silu(R, gate) * up)


def gelu_fast(R: Rounder, x):
    """nn.GELU(approx="fast") = x * mx.sigmoid(1.702 * x)"""
    c = R.scalar(1.702)
    return R.r(x * R.r(torch.sigmoid(R.r(c * x))))


def gelu_exact(R: Rounder, x):
    """nn.GELU() = x * (1 + mx.erf(x / math.sqrt(2))) / 2"""
    s2 = R.scalar(math.sqrt(2.0))
    a = R.r(x / s2)
    b = R.r(torch.erf(a))
    c = R.r(1.0 + b)
    d = R.r(x * c)
    return R.r(d / 2.0)


def gelu_tanh(R: Rounder, x):
    """nn.GELU(approx="precise") = 0.5 * x * (1 + tanh(sqrt(2 / pi) * (x + 0.044715 * x ** 3))) (mlx nn.gelu_approx);
    every primitive rounds to the activation dtype, python scalars are weak-typed to it."""
    c0, c1 = R.scalar(0.044715), R.scalar(math.sqrt(2.0 / math.pi))
    x3 = R.r(torch.pow(x, 3))     # mx.power(x, 3): one primitive, one rounding
    t = R.r(x + R.r(c0 * x3))
    t = R.r(torch.tanh(R.r(c1 * t)))
    return R.r(R.r(0.5 * x) * R.r(1.0 + t))


# ---------------------------------------------------------------------------
# mx.fast.scaled_dot_product_attention on the CPU device = fallback graph
# (mlx/fast.cpp): q = q * array(scale, q.dtype); GQA by reshaping q to
# (B, n_kv, n_rep, L, D); scores = q @ k^T (act dtype); causal mask is
# bottom-right aligned; softmax(precise=True) computes in fp32 and casts;
# out = scores @ v.
#   q,k,v: (B, H, L, D) / (B, Hkv, S, D) -> (B, H, L, D)
# ---------------------------------------------------------------------------
def sdpa(R: Rounder, q, k, v, scale: float, causal: bool, mask=None):
    """`mask` (bool, broadcastable to (B, 1, L, S), True = attend) is the array-mask form used by the
    batched path (left-padded rows, models/cache.py:24-42 with `left_padding`): the fallback graph
    replaces masked scores by the most negative FINITE value of the score dtype, so a fully masked
    (padding) query row yields a uniform, finite distribution instead of NaN."""
    B, H, L, D = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    rep = H // Hkv
    qs = R.r(q * R.scalar(scale))
    qs = qs.reshape(B, Hkv, rep, L, D)
    scores = R.r(qs @ k[:, :, None].transpose(-1, -2))  # (B,Hkv,rep,L,S)
    if mask is not None:
        m = torch.as_tensor(mask, dtype=torch.bool)
        while m.ndim < 4:
            m = m[None]
        m = m[:, :, None]  # (B|1, 1, 1, L, S)
        scores = torch.where(m, scores, torch.tensor(float(torch.finfo(R.dtype).min)))
    elif causal and L > 1:
        qi = torch.arange(S - L, S)[:, None]
        ki = torch.arange(S)[None, :]
        scores = torch.where(qi >= ki, scores, torch.tensor(float("-inf")))
    p = R.r(torch.softmax(scores, dim=-1))
    out = R.r(p @ v[:, :, None])
    return out.reshape(B, H, L, D)


def logsumexp(R: Rounder, x):
    """mx.logsumexp: fp32 inside, result in the input dtype."""
    return R.r(torch.logsumexp(x, dim=-1, keepdim=True))


def argmax_lowest(x):
    """mx.argmax: first (lowest-index) maximum."""
    m = x.max(dim=-1, keepdim=True).values
    idx = torch.arange(x.shape[-1]).expand_as(x)
    big = torch.full_like(idx, x.shape[-1])
    return torch.where(x == m, idx, big).min(dim=-1).values
