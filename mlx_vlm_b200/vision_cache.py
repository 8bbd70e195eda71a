"""`VisionFeatureCache` of the reference (mlx_vlm/vision_cache.py:15-79): an LRU of projected image
features keyed by the image source, consulted by `stream_generate` (generate/dispatch.py:800-809)
so that a multi-turn conversation about the same image skips the vision tower.  Values are whatever
`model.encode_image` returns (device tensors here) and are handed back to the model as
`cached_image_features` (consumed in `Model.get_input_embeddings`, qwen2_vl.py:50-57)."""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Any, Optional


class VisionFeatureCache:
    def __init__(self, max_size: int = 20):
        self.max_size = max_size
        self._cache: "OrderedDict[str, Any]" = OrderedDict()

    def _make_key(self, image_source: Any) -> str:
        """str / path -> itself; list -> keys joined by '|'; objects with `tobytes` (PIL images,
        arrays) -> 'pil:' + first 16 hex digits of their sha256; anything else -> its identity."""
        if isinstance(image_source, str):
            return image_source
        if isinstance(image_source, list):
            return "|".join(self._make_key(x) for x in image_source)
        if hasattr(image_source, "tobytes"):
            return "pil:" + hashlib.sha256(image_source.tobytes()).hexdigest()[:16]
        return f"obj:{id(image_source)}"

    def get(self, image_source: Any) -> Optional[Any]:
        key = self._make_key(image_source)
        if key not in self._cache:
            return None
        self._cache.move_to_end(key)
        return self._cache[key]

    def put(self, image_source: Any, features: Any) -> None:
        key = self._make_key(image_source)
        if key in self._cache:
            self._cache.move_to_end(key)
        elif len(self._cache) >= self.max_size:
            self._cache.popitem(last=False)  # least recently used
        self._cache[key] = features

    def clear(self) -> None:
        self._cache.clear()

    def __len__(self) -> int:
        return len(self._cache)

    def __contains__(self, image_source: Any) -> bool:
        return self._make_key(image_source) in self._cache
