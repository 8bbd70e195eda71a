"""ORACLE — TEST INFRASTRUCTURE ONLY. Not imported by the product path.

CPU restatement of the reference's LLaVA-Next (LLaVA-1.6) generate path, a sibling of LLaVA-1.5 that shares its
tower, projector and language model:
  models/llava_next/llava_next.py:47-96    get_input_embeddings: `pixel_values[0]` = the crops of ONE image
                                           (N, C, H, W) -> NHWC -> CLIP tower on the N crops -> hidden state
                                           `vision_feature_layer` -> class token dropped ("default") -> projector ->
                                           the `image_newline` vector broadcast to the features' shape and
                                           concatenated ALONG THE CROP AXIS (N crops become 2N blocks of P rows)
  models/llava_next/llava_next.py:98-121   merge: every <image> token of input_ids is REPLACED by one block of P
                                           rows (the sequence grows); zip(text segments, blocks) silently drops
                                           the blocks that have no <image> token left
Both are pinned by executing the reference's own source (tests/golden/make_llava_next_golden.py).  The tower,
projector and language model are oracle/llava.py's (fp32 tower with bf16-valued weights, one rounding at the merge)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from . import llava as L
from . import mlx_semantics as S
from . import qwen2vl as Q
from .mlx_semantics import Rounder

LlavaNextCfg = L.LlavaCfg


def init_weights(cfg, seed: int = 0):
    W = L.init_weights(cfg, seed)
    g = torch.Generator().manual_seed(seed * 7919 + 13)
    H = cfg.text.hidden_size
    W["image_newline"] = (torch.randn(H, generator=g) / H ** 0.5).to(torch.bfloat16).to(torch.float32)
    return W


def image_blocks(cfg, W, pixel_values_1nchw, R: Rounder) -> torch.Tensor:
    """(1, N, C, H, W) -> (2N, P, hidden): N projected crops followed by N blocks filled with `image_newline`"""
    pv = torch.as_tensor(np.asarray(pixel_values_1nchw), dtype=torch.float32)[0].permute(0, 2, 3, 1)
    states = L.clip_forward(cfg, W, pv, R)
    sel = states[cfg.vision_feature_layer]
    if cfg.vision_feature_select_strategy == "default":
        sel = sel[:, 1:]
    elif cfg.vision_feature_select_strategy != "full":
        raise ValueError(f"Unexpected feature selection strategy: {cfg.vision_feature_select_strategy}")
    y = S.linear(R, sel, W["multi_modal_projector.linear_1.weight"], W["multi_modal_projector.linear_1.bias"])
    y = S.gelu_exact(R, y)
    f = S.linear(R, y, W["multi_modal_projector.linear_2.weight"], W["multi_modal_projector.linear_2.bias"])
    nl = W["image_newline"].to(f.dtype)[None, None, :].expand_as(f)
    return torch.cat([f, nl], dim=0)


def merge(cfg, blocks: torch.Tensor, inputs_embeds: torch.Tensor, input_ids) -> torch.Tensor:
    """llava_next.py:98-121.  blocks (n, P, H); inputs_embeds (1, T, H)"""
    ids = np.asarray(input_ids)
    positions = np.where(ids == cfg.image_token_index)[1].tolist()
    segs, start = [], 0
    for p in positions:
        segs.append(inputs_embeds[:, start:p])
        start = p + 1
    chunks = [blocks[i:i + 1] for i in range(blocks.shape[0])]
    out: List[torch.Tensor] = [v for pair in zip(segs, chunks) for v in pair]
    out.append(inputs_embeds[:, start:])
    return torch.cat(out, dim=1)


def expanded_ids(cfg, input_ids, n_blocks: int, rows_per_block: int):
    """the positions of the merged sequence that hold image rows, as (ids with every used <image> replaced by
    `rows_per_block` copies, number of blocks used) — what the product's host logic must reproduce"""
    ids = np.asarray(input_ids)[0].tolist()
    out, used = [], 0
    for t in ids:
        if t == cfg.image_token_index:
            if used < n_blocks:
                out += [cfg.image_token_index] * rows_per_block
                used += 1
            # an <image> token without a block left: zip() ends there — see merge(): the token is dropped and
            # NOTHING after the last paired segment survives except the tail after the LAST <image> token
        else:
            out.append(t)
    return out, used


def get_input_embeddings(cfg, W, input_ids, pixel_values_1nchw, R: Rounder, vision_dtype: str = "f32"):
    ids = torch.as_tensor(np.asarray(input_ids), dtype=torch.long)
    embeds = W["language_model.model.embed_tokens.weight"][ids]
    if pixel_values_1nchw is None:
        return embeds, None
    blocks = R.r(image_blocks(cfg, W, pixel_values_1nchw, Rounder(vision_dtype)))   # astype(inputs_embeds.dtype)
    return merge(cfg, blocks, embeds, input_ids), blocks


def greedy_generate(cfg, W, input_ids, pixel_values_1nchw, max_tokens: int, dtype: str = "bf16",
                    vision_dtype: str = "f32"):
    R = Rounder(dtype)
    qc, W2 = L._as_qwen(cfg, W)
    embeds, blocks = get_input_embeddings(cfg, W, input_ids, pixel_values_1nchw, R, vision_dtype)
    T = embeds.shape[1]
    cache = [Q.OracleKVCache() for _ in range(cfg.text.num_hidden_layers)]
    hidden = Q.lm_layers_forward(qc, W2, embeds, L._positions(0, T), cache, R)
    logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
    out_logits, toks = [logits], []
    for n in range(max_tokens):
        y = S.argmax_lowest(Q.logprobs_from_logits(R, logits))
        toks.append(int(y[0]))
        if n == max_tokens - 1:
            break
        e = W["language_model.model.embed_tokens.weight"][y][:, None, :]
        hidden = Q.lm_layers_forward(qc, W2, e, L._positions(cache[0].offset, 1), cache, R)
        logits = Q.lm_head(qc, W2, hidden[:, -1, :], R)
        out_logits.append(logits)
    return {"tokens": toks, "logits": out_logits, "image_blocks": blocks, "inputs_embeds": embeds}
