"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's LOCK-STEP batched generation: what a
batched decode kernel has to reproduce.
  generate/ar.py:675-722   `_merge_prefill_prompt_kwargs`: per-row `inputs_embeds` (each row embedded on
                           its own, vision included) LEFT-padded with zeros to the longest prompt;
                           per-row M-RoPE position ids left-padded with zeros and concatenated on the
                           batch axis -> (3, B, Lmax); per-row `rope_deltas` -> (B, 1)
  models/cache.py:972-1201 `BatchKVCache(left_padding)`: one write index `_idx` for all rows, row b's
                           first `left_padding[b]` slots are dead, `offset[b] = _idx - left_padding[b]`
  models/cache.py:24-42    mask: key j visible to query i of row b iff j <= i and j >= left_padding[b]
  models/qwen2_vl/language.py:404-518 decode positions: offset[b] + rope_delta[b] on all three axes

The round-1 product time-multiplexes rows over the batch-1 engine, which is equivalent BY CONSTRUCTION
to running every row alone; this module states the lock-step formulation and
tests/test_oracle_batching.py shows it gives the same tokens / logits as the rows alone, i.e. it is the
acceptance oracle for the round-2 batched kernel (padding positions are dead weight, not semantics).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder


def left_padded_mask(N: int, offset: int, left_padding: Sequence[int]) -> torch.Tensor:
    """(B, 1, N, offset+N) bool, True = attend (create_causal_mask with left_padding)."""
    r = torch.arange(offset + N)[None, :]
    l = torch.arange(offset, offset + N)[:, None]
    m = (l >= r)[None, None]
    lp = torch.as_tensor(list(left_padding)).reshape(-1, 1, 1, 1)
    return m & (r[None, None] >= lp)


def batched_greedy_generate(cfg: Q.Cfg, W, requests: List[dict], max_tokens: int, dtype: str = "bf16"):
    """requests: dicts with input_ids (1, T_b) and optionally pixel_values / image_grid_thw.
    Returns per-row tokens (B, max_tokens), per-step logits [(B, V)], left_padding."""
    R = Rounder(dtype)
    t = cfg.text
    B = len(requests)
    rows = [Q.get_input_embeddings(cfg, W, np.asarray(r["input_ids"], dtype=np.int64),
                                   r.get("pixel_values"), r.get("image_grid_thw"), R) for r in requests]
    lens = [int(e.shape[1]) for e, _, _, _ in rows]
    Lmax = max(lens)
    pad = [Lmax - n for n in lens]
    H = rows[0][0].shape[-1]
    embeds = torch.zeros(B, Lmax, H)
    pos = np.zeros((3, B, Lmax), dtype=np.int64)
    deltas = np.zeros((B, 1), dtype=np.int64)
    for b, ((e, _, p, d), n) in enumerate(zip(rows, lens)):
        embeds[b, Lmax - n:] = e[0]
        p = np.asarray(p)
        if p.ndim == 2:
            p = np.broadcast_to(p[None], (3,) + p.shape)
        pos[:, b, Lmax - n:] = p[:, 0]
        deltas[b, 0] = int(np.asarray(d).reshape(-1)[0])
    cache = [Q.OracleKVCache() for _ in range(t.num_hidden_layers)]
    hidden = Q.lm_layers_forward(cfg, W, embeds, pos, cache, R, mask=left_padded_mask(Lmax, 0, pad))
    logits = Q.lm_head(cfg, W, hidden[:, -1, :], R)
    toks, all_logits = [], []
    for n in range(max_tokens):
        all_logits.append(logits.clone())
        y = S.argmax_lowest(Q.logprobs_from_logits(R, logits))
        toks.append(y.clone())
        if n == max_tokens - 1:
            break
        e = W["language_model.model.embed_tokens.weight"][y][:, None, :]
        idx = cache[0].offset                                  # shared write index
        row_off = np.asarray([idx - p for p in pad])           # BatchKVCache.offset
        p1 = (row_off + deltas[:, 0])[None, :, None]
        hidden = Q.lm_layers_forward(cfg, W, e, np.broadcast_to(p1, (3, B, 1)).copy(), cache, R,
                                     mask=left_padded_mask(1, idx, pad))
        logits = Q.lm_head(cfg, W, hidden[:, -1, :], R)
    return {"tokens": torch.stack(toks, 1), "logits": all_logits, "left_padding": pad, "rope_deltas": deltas}
