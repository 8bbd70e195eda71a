"""Multi-GPU plumbing of the generate path: requests are independent
units, so the design is one full model replica per GPU (one process per GPU,
`torch.distributed`), a single weight broadcast from rank 0 at start-up (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests), a host-side request router,
and NO per-step collective.  Timing results are max-reduced over ranks.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def broadcast_weights(weights: Dict[str, torch.Tensor], src: int = 0, group=None) -> int:
    """Broadcast every tensor (sorted by name, in place) from `src`; returns bytes moved."""
    n = 0
    for name in sorted(weights):
        t = weights[name]
        dist.broadcast(t, src=src, group=group)
        n += t.numel() * t.element_size()
    return n


def shard_requests(n_requests: int, world_size: int, rank: int) -> List[int]:
    """Round-robin request ids owned by `rank` (every id owned by exactly one rank)."""
    return list(range(rank, n_requests, world_size))


def least_loaded(loads: Sequence[int]) -> int:
    """Router policy for a live queue: the replica with the fewest active rows."""
    best = 0
    for i, v in enumerate(loads):
        if v < loads[best]:
            best = i
    return best


def max_over_ranks(values: Sequence[float], device=None, group=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t.tolist()
