"""Image-feature cache for multi-turn conversations.

Host mirror of the reference's `VisionFeatureCache` API (mlx_vlm/vision_cache.py:15-79): `get` /
`put` / `clear` / `len` / `in`, least-recently-used eviction at `max_size`, and the same key rule
(pinned by tests/test_batch_host.py::test_vision_feature_cache_lru_and_keys).  `stream_generate`
consults it (generate/dispatch.py:800-809); a hit travels to the model as `cached_image_features`
(device tensors here) and skips the vision tower.
"""
from __future__ import annotations

import hashlib
from typing import Any, Dict, Optional


def image_key(source: Any) -> str:
    """Identity of an image source: a path / URL is its own key; a list joins its members' keys with
    '|'; anything exposing `tobytes()` (PIL image, ndarray) is keyed by 'pil:' + the first 16 hex
    digits of the sha256 of those bytes; other objects by 'obj:' + id()."""
    if isinstance(source, str):
        return source
    if isinstance(source, list):
        return "|".join(map(image_key, source))
    raw = getattr(source, "tobytes", None)
    if callable(raw):
        return "pil:" + hashlib.sha256(raw()).hexdigest()[:16]
    return "obj:%d" % id(source)


class VisionFeatureCache:
    """Python dicts keep insertion order: the first key is always the least recently used one, and a
    touched entry is re-inserted at the end."""

    def __init__(self, max_size: int = 20):
        self.max_size = max_size
        self._entries: Dict[str, Any] = {}

    _make_key = staticmethod(image_key)

    def _touch(self, key: str, value: Any) -> None:
        self._entries.pop(key, None)
        self._entries[key] = value

    def get(self, image_source: Any) -> Optional[Any]:
        key = image_key(image_source)
        if key not in self._entries:
            return None
        value = self._entries[key]
        self._touch(key, value)
        return value

    def put(self, image_source: Any, features: Any) -> None:
        key = image_key(image_source)
        if key not in self._entries and len(self._entries) >= self.max_size:
            del self._entries[next(iter(self._entries))]
        self._touch(key, features)

    def clear(self) -> None:
        self._entries = {}

    def __len__(self) -> int:
        return len(self._entries)

    def __contains__(self, image_source: Any) -> bool:
        return image_key(image_source) in self._entries
