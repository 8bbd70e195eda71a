"""Multi-GPU plumbing of the generate path: requests are independent
units, so the design is one full model replica per GPU (one process per GPU,
`torch.distributed`), a single weight broadcast from rank 0 at start-up (NCCL over
NVLink / NVSwitch on the GPU box, gloo in the CPU tests), a host-side request router,
and NO per-step collective.  Timing results are max-reduced over ranks.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

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


def broadcast_packed(flat: torch.Tensor, src: int = 0, group=None) -> int:
    """THE collective of the design: every weight of a replica is a view of one flat buffer
    (`Model.packed_weights`), so the start-up is ONE broadcast (NCCL over NVLink / NVSwitch) instead
    of one per tensor.  Returns bytes moved."""
    dist.broadcast(flat, src=src, group=group)
    return flat.numel() * flat.element_size()


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


def generate_sharded(make_generator: Callable[[], Any], prompts: Sequence[Sequence[int]],
                     max_tokens, prompt_kwargs: Optional[List[dict]] = None, group=None) -> List[List[int]]:
    """Config C5 on N replicas: request i runs on rank i % world (round-robin router), every rank drives
    its own `BatchGenerator` (`make_generator()` builds it around the rank's replica) to completion,
    and the generated token ids come back on EVERY rank in request order.  The only communication is
    the final gather of python lists — nothing per step."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    mine = shard_requests(len(prompts), world, rank)
    if isinstance(max_tokens, int):
        max_tokens = [max_tokens] * len(prompts)
    out: Dict[int, List[int]] = {}
    if mine:
        gen = make_generator()
        kws = [dict((prompt_kwargs or [{}] * len(prompts))[i]) for i in mine]
        uids = gen.insert([list(prompts[i]) for i in mine], [max_tokens[i] for i in mine], kws)
        by_uid = {u: i for u, i in zip(uids, mine)}
        for i in mine:
            out[i] = []
        while gen.has_work:
            _, responses = gen.next()
            for r in responses:
                out[by_uid[r.uid]].append(r.token)
    if world > 1:
        parts: List[Optional[Dict[int, List[int]]]] = [None] * world
        dist.all_gather_object(parts, out, group=group)
        out = {k: v for part in parts for k, v in part.items()}
    return [out[i] for i in range(len(prompts))]
