#!/usr/bin/env python
"""Golden vectors for Qwen2.5-VL: the reference's own `VisionModel.get_window_index`, `rot_pos_emb` and the
integer bookkeeping of `VisionModel.__call__` (mlx_vlm/models/qwen2_5_vl/vision.py: window permutation of the merge
units, de-duplicated window boundaries, per-frame boundaries, which blocks see which, the reverse permutation after the
merger) are extracted with `ast` and EXECUTED over the numpy stand-in for mlx.core of make_golden.py; the blocks and the
merger are stubs that record what they are called with.
Writes tests/golden/qwen2_5_vl_golden.json.   usage: python tests/golden/make_qwen2_5_vl_golden.py"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load, make_mx, tolist  # noqa: E402

OUT = os.path.join(HERE, "qwen2_5_vl_golden.json")


def main():
    from typing import Optional
    mx = make_mx()
    mx.arange = lambda *a, dtype=None: np.arange(*a, dtype=dtype)
    mx.array = lambda x, dtype=None: np.array(x, dtype=dtype)
    ns = {"mx": mx, "np": np, "Optional": Optional}
    gwi, w1 = load(ns, "models/qwen2_5_vl/vision.py", "get_window_index", "VisionModel")
    rpe, w2 = load(ns, "models/qwen2_5_vl/vision.py", "rot_pos_emb", "VisionModel")
    call, w3 = load(ns, "models/qwen2_5_vl/vision.py", "__call__", "VisionModel")
    golden = {"_about": "reference qwen2_5_vl VisionModel integer logic executed over a numpy stand-in",
              "provenance": {"get_window_index": w1, "rot_pos_emb": w2, "__call__": w3}, "cases": []}
    cases = [
        (112, 14, 2, [[1, 4, 4]]),              # one window exactly
        (112, 14, 2, [[1, 16, 16]]),            # 2 x 2 full windows (the reference still pads one more window row/col)
        (112, 14, 2, [[1, 24, 24]]),            # 336 x 336: 12 x 12 merge units = 3 x 3 windows + empty padding windows
        (112, 14, 2, [[1, 12, 20]]),            # ragged: partial windows on both edges
        (112, 14, 2, [[1, 6, 10], [1, 20, 8]]), # two images
        (112, 14, 2, [[2, 8, 12]]),             # two temporal frames
        (56, 14, 2, [[1, 10, 6], [1, 4, 4]]),   # window of 2 x 2 merge units
    ]
    for window, patch, ms, grid in cases:
        g = np.asarray(grid, dtype=np.int64)
        self = types.SimpleNamespace(window_size=window, patch_size=patch, spatial_merge_size=ms,
                                     spatial_merge_unit=ms * ms)
        widx, cuw = gwi(self, g)
        # rot_pos_emb with a table whose row r is [r, r + 0.5]: the output encodes (h id, w id) per patch
        self.rotary_pos_emb = lambda n: np.stack([np.arange(int(n), dtype=np.float32),
                                                  np.arange(int(n), dtype=np.float32) + 0.5], -1)
        rot = rpe(self, g)
        seq = int((g[:, 0] * g[:, 1] * g[:, 2]).sum())
        seen = []

        def block(h, cu_seqlens=None, rotary_pos_emb=None):
            seen.append({"cu": tolist(cu_seqlens), "order": tolist(h[:, 0].astype(np.int64)),
                         "rot": tolist(rotary_pos_emb)})
            return h
        self.patch_embed = lambda x: x
        self.rot_pos_emb = lambda gg: rot
        self.get_window_index = lambda gg: gwi(self, gg)
        self.fullatt_block_indexes = [1]
        self.blocks = [block, block, block]
        self.merger = lambda h: h.reshape(-1, ms * ms)        # one output row per merge unit: its 4 patch ids
        x = np.arange(seq, dtype=np.float32)[:, None]
        out = call(self, x, g)
        golden["cases"].append({
            "window_size": window, "patch_size": patch, "merge": ms, "grid_thw": grid,
            "window_index": tolist(widx), "cu_window_seqlens_raw": tolist(cuw),
            "rot_hw": tolist(rot), "blocks": [{"cu": b["cu"]} for b in seen], "order": seen[0]["order"],
            "rot_in_window_order": seen[0]["rot"], "out": tolist(out.astype(np.int64))})
    with open(OUT, "w") as f:
        json.dump(golden, f)
    print("wrote", OUT, len(golden["cases"]))
    for c in golden["cases"]:
        print(c["grid_thw"], "windows", len(c["blocks"][0]["cu"]) - 1, "raw", len(c["cu_window_seqlens_raw"]) - 1,
              "full", c["blocks"][1]["cu"])


if __name__ == "__main__":
    main()
