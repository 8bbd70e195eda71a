#!/usr/bin/env python
"""Golden vectors for Idefics3 / SmolVLM: the reference's own `Idefics3Connector.pixel_shuffle`
(mlx_vlm/models/idefics3/idefics3.py) and `VisionEmbeddings.__call__` (idefics3/vision.py: bucketed position ids —
NOT Idefics2's digitize quirk —, ids written to the first `valid` positions, position embeddings zeroed on padding
patches) are extracted with `ast` and EXECUTED over the numpy stand-in for mlx.core of make_golden.py.
Writes tests/golden/idefics3_golden.json.   usage: python tests/golden/make_idefics3_golden.py"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load, make_mx, tolist  # noqa: E402

OUT = os.path.join(HERE, "idefics3_golden.json")


def main():
    mx = make_mx()
    mx.tile = np.tile
    mx.flatten = lambda a, start_axis=0, end_axis=-1: np.reshape(
        a, a.shape[:start_axis] + (-1,) + (a.shape[end_axis + 1:] if end_axis != -1 and end_axis + 1 < a.ndim else ()))
    def arange(*a, dtype=None):
        """integers: numpy; python floats: float32 with element i = start + i * step, product rounded to float32 — mlx's
        Metal `arange` kernel (numpy would compute the boundaries in float64 and break the ties of the `>=` below
        differently; see oracle/idefics3.py::mlx_arange_f32)"""
        if dtype is None and any(isinstance(x, float) for x in a):
            start, stop, step = a
            n = max(int(np.ceil((stop - start) / step)), 0)
            s0 = np.float32(start)
            st = np.float32(np.float32(start + step) - s0)
            return np.array([np.float32(s0 + np.float32(np.float32(i) * st)) for i in range(n)], dtype=np.float32)
        return np.arange(*a, dtype=dtype)
    mx.arange = arange
    mx.clip = lambda a, a_min=None, a_max=None: np.clip(a, np.float32(a_min), np.float32(a_max)).astype(np.float32)
    mx.zeros = lambda shape, dtype=None: np.zeros(shape, dtype=dtype)
    ns = {"mx": mx, "np": np}
    shuffle, w1 = load(ns, "models/idefics3/idefics3.py", "pixel_shuffle", "Idefics3Connector")
    embed, w2 = load(ns, "models/idefics3/vision.py", "__call__", "VisionEmbeddings")
    golden = {"_about": "reference idefics3 functions executed over a numpy stand-in (make_idefics3_golden.py)",
              "provenance": {"pixel_shuffle": w1, "VisionEmbeddings.__call__": w2}, "pixel_shuffle": [], "embeddings": []}
    # ---- pixel shuffle: out[i, :] as indices of the source tokens
    for side, s, E in ((4, 2, 3), (6, 2, 2), (6, 3, 1), (8, 4, 2)):
        seq = side * side
        x = (np.arange(seq, dtype=np.float32)[:, None] * 100 + np.arange(E, dtype=np.float32)[None, :])[None]
        out = shuffle(types.SimpleNamespace(), mx.array(x), s)
        golden["pixel_shuffle"].append({"side": side, "scale": s, "E": E, "out_shape": list(out.shape),
                                        "out": tolist(out[0])})
    # ---- vision embeddings: patch embedding stubbed to 0 -> output = position embedding rows * mask
    E, side = 2, 5            # 5 x 5 position grid
    table = (np.arange(side * side, dtype=np.float32)[:, None] * 10 + np.arange(E, dtype=np.float32)[None, :] + 1.0)
    for (gh, gw, vh, vw) in ((5, 5, 5, 5), (4, 3, 4, 3), (4, 4, 2, 3), (3, 5, 3, 1), (2, 2, 0, 0)):
        mask = np.zeros((1, gh, gw), dtype=bool)
        mask[0, :vh, :vw] = True
        self = types.SimpleNamespace(
            patch_embedding=lambda x, gh=gh, gw=gw: np.zeros((x.shape[0], gh, gw, E), dtype=np.float32),
            position_embedding=lambda ids: table[np.asarray(ids)], num_patches_per_side=side)
        x = np.zeros((1, gh * 14, gw * 14, 3), dtype=np.float32)
        out = embed(self, mx.array(x), mx.array(mask))
        golden["embeddings"].append({"grid": [gh, gw], "valid": [vh, vw], "side": side, "E": E, "table": tolist(table),
                                     "patch_mask": mask[0].astype(int).tolist(), "out": tolist(out[0])})
    # grids where the float32 rounding of the boundaries decides ties: the Idefics3 (26) and SmolVLM (27) position grids
    golden["position_ids"] = []
    for side, vh, vw in ((26, 26, 26), (26, 13, 26), (27, 27, 27), (27, 27, 9), (6, 6, 6)):
        ids_table = np.arange(side * side, dtype=np.float32)[:, None]
        mask = np.zeros((1, side, side), dtype=bool)
        mask[0, :vh, :vw] = True
        self = types.SimpleNamespace(
            patch_embedding=lambda x, side=side: np.zeros((x.shape[0], side, side, 1), dtype=np.float32),
            position_embedding=lambda ids: ids_table[np.asarray(ids)], num_patches_per_side=side)
        out = embed(self, mx.array(np.zeros((1, side * 14, side * 14, 3), dtype=np.float32)), mx.array(mask))
        golden["position_ids"].append({"side": side, "valid": [vh, vw],
                                       "ids_times_mask": np.asarray(out[0, :, 0]).astype(np.int64).tolist()})
    x = np.zeros((1, 3 * 14, 4 * 14, 3), dtype=np.float32)
    self = types.SimpleNamespace(patch_embedding=lambda x: np.zeros((1, 3, 4, E), dtype=np.float32),
                                 position_embedding=lambda ids: table[np.asarray(ids)], num_patches_per_side=side)
    out = embed(self, mx.array(x), None)
    golden["embeddings"].append({"grid": [3, 4], "valid": None, "side": side, "E": E, "table": tolist(table),
                                 "patch_mask": None, "out": tolist(out[0])})
    with open(OUT, "w") as f:
        json.dump(golden, f)
    print("wrote", OUT, len(golden["pixel_shuffle"]), len(golden["embeddings"]))


if __name__ == "__main__":
    main()
