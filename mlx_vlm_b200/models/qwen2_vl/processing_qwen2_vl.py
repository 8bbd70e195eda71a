"""Host preprocessing for Qwen2-VL (numpy, torch-free), following the reference's
mlx_vlm/models/qwen3_vl/processing_qwen3_vl.py:182-205 (`smart_resize`), :208-227
(`_to_numpy_image`), :302-354 (`_process_one`) and
models/qwen2_vl/processing_qwen2_vl.py:62-127 (placeholder expansion).

`SyntheticProcessor` stands in for a checkpoint's tokenizer when no tokenizer
files exist (no network in the build/bench environment): it produces
deterministic token ids of a requested length but runs the REAL image pipeline.
"""
from __future__ import annotations

import math
import zlib
from typing import List, Optional, Sequence

import numpy as np

OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def smart_resize(height: int, width: int, factor: int = 28, min_pixels: int = 56 * 56,
                 max_pixels: int = 14 * 14 * 4 * 1280):
    if max(height, width) / min(height, width) > 200:
        raise ValueError("absolute aspect ratio must be smaller than 200, got "
                         f"{max(height, width) / min(height, width)}")
    h_bar = round(height / factor) * factor
    w_bar = round(width / factor) * factor
    if h_bar * w_bar > max_pixels:
        beta = math.sqrt((height * width) / max_pixels)
        h_bar = max(factor, math.floor(height / beta / factor) * factor)
        w_bar = max(factor, math.floor(width / beta / factor) * factor)
    elif h_bar * w_bar < min_pixels:
        beta = math.sqrt(min_pixels / (height * width))
        h_bar = math.ceil(height * beta / factor) * factor
        w_bar = math.ceil(width * beta / factor) * factor
    return h_bar, w_bar


def to_numpy_image(img) -> np.ndarray:
    """PIL.Image / path / ndarray -> (C,H,W) uint8 (or float) array."""
    if isinstance(img, str):
        from PIL import Image
        img = Image.open(img)
    if hasattr(img, "convert"):
        arr = np.array(img.convert("RGB"))
    else:
        arr = np.asarray(img)
    if arr.ndim == 2:
        arr = np.stack([arr] * 3, axis=-1)
    if arr.ndim == 3 and arr.shape[-1] in (1, 3, 4):
        arr = np.transpose(arr, (2, 0, 1))
    if arr.shape[0] == 4:
        arr = arr[:3]
    return arr


class Qwen2VLImageProcessor:
    model_input_names = ["pixel_values", "image_grid_thw"]

    def __init__(self, patch_size=14, temporal_patch_size=2, merge_size=2, min_pixels=56 * 56,
                 max_pixels=14 * 14 * 4 * 1280, do_rescale=True, rescale_factor=1 / 255.0,
                 do_normalize=True, image_mean=None, image_std=None, **_):
        self.patch_size, self.temporal_patch_size, self.merge_size = patch_size, temporal_patch_size, merge_size
        self.min_pixels, self.max_pixels = min_pixels, max_pixels
        self.do_rescale, self.rescale_factor, self.do_normalize = do_rescale, rescale_factor, do_normalize
        self.image_mean = list(image_mean or [0.5, 0.5, 0.5])
        self.image_std = list(image_std or [0.5, 0.5, 0.5])

    def num_image_tokens(self, height: int, width: int) -> int:
        h, w = smart_resize(height, width, self.patch_size * self.merge_size, self.min_pixels,
                            self.max_pixels)
        return (h // self.patch_size) * (w // self.patch_size) // (self.merge_size ** 2)

    def _process_one(self, image: np.ndarray):
        C, H, W = image.shape
        rh, rw = smart_resize(H, W, self.patch_size * self.merge_size, self.min_pixels,
                              self.max_pixels)
        if (rh, rw) != (H, W):  # unchanged sizes take no resample (:164-170)
            from PIL import Image
            arr = np.transpose(image, (1, 2, 0))
            if arr.dtype in (np.float32, np.float64):
                arr = (arr * 255).clip(0, 255).astype(np.uint8)
            pil = Image.fromarray(arr).resize((rw, rh), resample=Image.BICUBIC)
            frame = np.transpose(np.array(pil), (2, 0, 1))
        else:
            frame = image
        img = frame.astype(np.float32)
        if self.do_rescale and image.dtype == np.uint8:
            img = img * np.float32(self.rescale_factor)
        if self.do_normalize:
            mean = np.array(self.image_mean, dtype=np.float32)[:, None, None]
            std = np.array(self.image_std, dtype=np.float32)[:, None, None]
            img = (img - mean) / std
        ps, tps, ms = self.patch_size, self.temporal_patch_size, self.merge_size
        gh, gw = rh // ps, rw // ps
        x = np.repeat(img[None, None], tps, axis=1)
        x = x.reshape(1, 1, tps, C, gh // ms, ms, ps, gw // ms, ms, ps)
        x = x.transpose(0, 1, 4, 7, 5, 8, 3, 2, 6, 9)
        return np.ascontiguousarray(x.reshape(gh * gw, C * tps * ps * ps)), [1, gh, gw]

    def __call__(self, images, **_):
        if not isinstance(images, (list, tuple)):
            images = [images]
        rows, grids = [], []
        for im in images:
            r, g = self._process_one(to_numpy_image(im))
            rows.append(r)
            grids.append(g)
        return {"pixel_values": np.concatenate(rows, axis=0),
                "image_grid_thw": np.asarray(grids, dtype=np.int64)}


class SyntheticTokenizer:
    """Deterministic stand-in tokenizer: ids are CRC32-derived, text ids < vocab_text."""

    def __init__(self, vocab_text: int, eos_token_id: Optional[int] = None):
        self.vocab_text = vocab_text
        self.eos_token_id = eos_token_id
        self.eos_token_ids = [] if eos_token_id is None else [eos_token_id]
        self.all_special_ids: List[int] = []
        self.stopping_criteria = None

    def encode(self, text: str, add_special_tokens: bool = False) -> List[int]:
        return [zlib.crc32(w.encode()) % self.vocab_text for w in text.split()]

    def decode(self, ids: Sequence[int], **_) -> str:
        return "".join(f"<{int(i)}>" for i in ids)


class SyntheticProcessor:
    """ProcessorLike (generate/types.py): `processor(text=, images=)` -> numpy dict with
    input_ids, attention_mask, pixel_values, image_grid_thw.  The prompt is
    `n_text_tokens` synthetic ids wrapped around <|vision_start|><|image_pad|>*N<|vision_end|>
    (N from the image grid, processing_qwen2_vl.py:93-105)."""

    def __init__(self, config, n_text_tokens: int = 128, image_mean=OPENAI_CLIP_MEAN,
                 image_std=OPENAI_CLIP_STD, seed: int = 0):
        self.config = config
        v = config.vision_config
        self.image_processor = Qwen2VLImageProcessor(
            patch_size=v.patch_size, temporal_patch_size=v.temporal_patch_size,
            merge_size=v.spatial_merge_size, image_mean=image_mean, image_std=image_std)
        self.n_text_tokens = n_text_tokens
        self.seed = seed
        hi = min(config.text_config.vocab_size, config.image_token_id) - 16
        self.tokenizer = SyntheticTokenizer(max(hi, 8))

    def __call__(self, text=None, images=None, **_):
        prompts = text if isinstance(text, (list, tuple)) else [text]
        out = {}
        grids = []
        if images is not None:
            im = self.image_processor(images)
            out.update(im)
            grids = im["image_grid_thw"].tolist()
        cfg = self.config
        m2 = cfg.vision_config.spatial_merge_size ** 2
        rows = []
        gi = 0
        for pi, p in enumerate(prompts):
            rng = np.random.default_rng(self.seed + zlib.crc32((p or "").encode()) + pi)
            ids = rng.integers(0, self.tokenizer.vocab_text, size=self.n_text_tokens).tolist()
            if grids:
                n_img = len(grids) if len(prompts) == 1 else 1
                vis = []
                for _ in range(n_img):
                    n = int(np.prod(grids[gi])) // m2
                    vis += [cfg.vision_start_token_id] + [cfg.image_token_id] * n + \
                           [getattr(cfg, "vision_end_token_id", cfg.vision_start_token_id + 1)]
                    gi += 1
                cut = min(4, len(ids) // 2)
                # keep exactly n_text_tokens text-side tokens incl. the start/end markers
                body = ids[: max(len(ids) - 2 * n_img, 0)]
                ids = body[:cut] + vis + body[cut:]
            rows.append(ids)
        L = max(len(r) for r in rows)
        ids = np.zeros((len(rows), L), dtype=np.int64)
        mask = np.zeros((len(rows), L), dtype=np.int64)
        for i, r in enumerate(rows):  # left padding (utils.py:1847-1891 padding_side="left")
            ids[i, L - len(r):] = r
            mask[i, L - len(r):] = 1
        out["input_ids"] = ids
        out["attention_mask"] = mask
        return out
