"""KV cache objects with the reference's Python-visible behaviour
(mlx_vlm/models/cache.py:337-439 `KVCache`, :45-70 `make_prompt_cache`).

Storage is B200-first: ONE device pool per request (`KVPool`, laid out
(n_layers, 2, batch, n_kv_heads, capacity, head_dim) bf16) that the CUDA engine
writes directly (K is rotated and appended inside the QKV kernels), instead of
per-layer arrays grown by concatenation.  Each layer's `KVCache` is a view of the
pool that keeps `offset / keys / values / state / trim / update_and_fetch /
is_trimmable / size / empty / nbytes` semantics; capacity still grows in
`step = 256` multiples (cache.py:338).
"""
from __future__ import annotations

from typing import Any, List, Optional

import torch


class KVPool:
    step = 256

    def __init__(self, n_layers: int, n_kv_heads: int, head_dim: int, device, batch: int = 1,
                 capacity: int = 0, dtype=torch.bfloat16):
        self.n_layers, self.n_kv, self.hd = n_layers, n_kv_heads, head_dim
        self.batch = batch
        self.device = device
        self.dtype = dtype
        self.buf: Optional[torch.Tensor] = None
        self.capacity = 0
        self.generation = 0  # bumped whenever the buffer moves (engine must re-bind)
        if capacity:
            self.reserve(capacity)

    def reserve(self, n_tokens: int, live_tokens: int = 0) -> bool:
        """Make room for n_tokens positions; returns True if the buffer moved."""
        if n_tokens <= self.capacity:
            return False
        cap = ((max(n_tokens, 2 * self.capacity) + self.step - 1) // self.step) * self.step
        new = torch.zeros((self.n_layers, 2, self.batch, self.n_kv, cap, self.hd),
                          dtype=self.dtype, device=self.device)
        if self.buf is not None and live_tokens > 0:
            new[..., :live_tokens, :].copy_(self.buf[..., :live_tokens, :])
        self.buf = new
        self.capacity = cap
        self.generation += 1
        return True


class _BaseCache:
    @property
    def state(self):
        return []

    def is_trimmable(self):
        return False

    def size(self):
        return 0

    def empty(self):
        raise NotImplementedError


class KVCache(_BaseCache):
    step = 256

    def __init__(self, pool: Optional[KVPool] = None, layer: int = 0):
        self._pool = pool
        self._layer = layer
        self.offset = 0
        self._own = None  # standalone storage when used without a pool (API parity)

    # --- reference attribute surface ---------------------------------------
    @property
    def keys(self):
        if self._pool is not None:
            return None if self._pool.buf is None else self._pool.buf[self._layer, 0]
        return None if self._own is None else self._own[0]

    @property
    def values(self):
        if self._pool is not None:
            return None if self._pool.buf is None else self._pool.buf[self._layer, 1]
        return None if self._own is None else self._own[1]

    def update_and_fetch(self, keys: torch.Tensor, values: torch.Tensor):
        """cache.py:345-367.  Generic (non-fused) append used by callers that
        bring their own K/V; the engine's kernels append in place instead."""
        prev = self.offset
        L = keys.shape[2]
        if self._pool is not None:
            self._pool.reserve(prev + L, live_tokens=prev)
            k, v = self.keys, self.values
        else:
            B, nkv, _, hd = keys.shape
            need = prev + L
            cap = 0 if self._own is None else self._own.shape[3]
            if need > cap:
                n_steps = (self.step + need - 1) // self.step
                new = torch.zeros((2, B, nkv, n_steps * self.step, hd), dtype=keys.dtype,
                                  device=keys.device)
                if self._own is not None:
                    new[:, :, :, :prev].copy_(self._own[:, :, :, :prev])
                self._own = new
            k, v = self._own[0], self._own[1]
        self.offset += L
        k[..., prev:self.offset, :] = keys
        v[..., prev:self.offset, :] = values
        return k[..., :self.offset, :], v[..., :self.offset, :]

    def size(self):
        return self.offset

    @property
    def state(self):
        return self.keys[..., :self.offset, :], self.values[..., :self.offset, :]

    @state.setter
    def state(self, v):
        k, vv = v
        self.offset = 0
        if self._pool is None:
            self._own = None
        self.update_and_fetch(k, vv)

    def is_trimmable(self):
        return True

    def trim(self, n):
        n = min(self.offset, n)
        self.offset -= n
        return n

    def extract(self, idx):
        cache = KVCache()
        if self.keys is None:
            if idx not in (0, -1):
                raise IndexError("KVCache row index out of range")
            return cache
        batch_size = int(self.keys.shape[0])
        if idx < 0:
            idx += batch_size
        if idx < 0 or idx >= batch_size:
            raise IndexError(f"KVCache row index {idx} out of range for batch size {batch_size}")
        cache.update_and_fetch(self.keys[idx:idx + 1, :, :self.offset, :].contiguous(),
                               self.values[idx:idx + 1, :, :self.offset, :].contiguous())
        return cache

    def make_mask(self, N: int, return_array: bool = False, window_size=None):
        # create_attention_mask (cache.py:73-83): "causal" for N>1, None for N==1
        if N == 1:
            return None
        return "causal"

    def empty(self):
        return self.keys is None or self.offset == 0 and self._pool is None and self._own is None

    @property
    def nbytes(self):
        k = self.keys
        if k is None:
            return 0
        return 2 * k.numel() * k.element_size()


def make_prompt_cache(model: Any, max_kv_size: Optional[int] = None) -> List[Any]:
    """cache.py:45-70: defer to model.make_cache() when present."""
    if max_kv_size is not None:
        raise NotImplementedError("RotatingKVCache (max_kv_size) is out of scope (SURVEY §2.1 #6)")
    if hasattr(model, "make_cache"):
        return model.make_cache()
    return [KVCache() for _ in range(len(model.layers))]
