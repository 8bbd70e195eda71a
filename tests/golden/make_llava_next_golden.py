#!/usr/bin/env python
"""Golden vectors for LLaVA-Next (a sibling of LLaVA-1.5 that shares its kernels): the reference's own
`Model.get_input_embeddings` and `_merge_input_ids_with_image_features` (mlx_vlm/models/llava_next/llava_next.py:47-127)
are extracted with `ast` from /root/reference and EXECUTED over the numpy stand-in for mlx.core of make_golden.py, with
stub sub-modules (tower: deterministic hidden states per crop, projector: x -> 2x + 1, embedding table: row lookup),
so that every indexing decision — `pixel_values[0]`, NCHW -> NHWC, feature layer, class-token drop, the "newline"
concatenated ALONG THE CROP AXIS, the zip() of text segments and crops that silently drops surplus crops — is the
reference's.  Writes tests/golden/llava_next_golden.json.   usage: python tests/golden/make_llava_next_golden.py"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load, make_mx, tolist  # noqa: E402

OUT = os.path.join(HERE, "llava_next_golden.json")


def main():
    mx = make_mx()
    mx.split = lambda a, n, axis=0: np.split(a, n, axis=axis)
    from typing import Optional
    feats_cls = type("InputEmbeddingsFeatures", (), {"__init__": lambda self, inputs_embeds=None: setattr(self, "inputs_embeds", inputs_embeds)})
    ns = {"mx": mx, "np": np, "Optional": Optional, "InputEmbeddingsFeatures": feats_cls}
    gie, w1 = load(ns, "models/llava_next/llava_next.py", "get_input_embeddings", "Model")
    merge, w2 = load(ns, "models/llava_next/llava_next.py", "_merge_input_ids_with_image_features", "Model")
    golden = {"_about": "reference llava_next functions executed over a numpy stand-in (make_llava_next_golden.py)",
              "provenance": {"get_input_embeddings": w1, "_merge_input_ids_with_image_features": w2}, "cases": []}
    H, P, V = 4, 3, 50                      # hidden, patches per crop (after the class-token drop), vocabulary
    table = (np.arange(V * H, dtype=np.float32).reshape(V, H) + 0.5)
    newline = np.asarray([9000.0, 9001.0, 9002.0, 9003.0], dtype=np.float32)

    def tower(x_nhwc, output_hidden_states=True):
        # hidden state l of crop n, token t (class token first), channel c = 100 (l + 1) + 1000 n + 10 t + c
        n = x_nhwc.shape[0]
        assert x_nhwc.shape[-1] == 3, "the reference hands NHWC crops to the tower"
        nn = np.arange(n, dtype=np.float32)[:, None, None]
        t = np.arange(P + 1, dtype=np.float32)[None, :, None]
        c = np.arange(H, dtype=np.float32)[None, None, :]
        return None, None, [100.0 * (l + 1) + 1000.0 * nn + 10.0 * t + c for l in range(4)]

    def run(ids, n_crops, strategy, layer, tag):
        ids = np.asarray(ids)
        self = types.SimpleNamespace(
            config=types.SimpleNamespace(image_token_index=32),
            vision_tower=tower, vision_feature_layer=layer, vision_feature_select_strategy=strategy,
            multi_modal_projector=lambda x: 2.0 * x + 1.0, image_newline=newline,
            language_model=types.SimpleNamespace(model=types.SimpleNamespace(embed_tokens=lambda i: table[np.asarray(i)])))
        self._merge_input_ids_with_image_features = lambda f, e, i: merge(self, f, e, i)
        pv = np.zeros((1, n_crops, 3, 2, 2), dtype=np.float32)
        try:
            out = gie(self, mx.array(ids), mx.array(pv))
            res, err = tolist(out.inputs_embeds), None
        except Exception as e:
            res, err = None, f"{type(e).__name__}: {e}"
        golden["cases"].append({"tag": tag, "input_ids": ids.tolist(), "n_crops": n_crops, "strategy": strategy,
                                "layer": layer, "hidden": H, "patches": P, "newline": newline.tolist(),
                                "output": res, "error": err})

    run([[5, 32, 6, 7]], 3, "default", -2, "one <image>, 3 crops: only crop 0 is inserted")
    run([[32, 5, 32, 6]], 2, "default", -2, "two <image> tokens, 2 crops: both crops, no newline rows")
    run([[5, 32, 6, 32, 7, 32]], 2, "default", -1, "three <image> tokens, 2 crops: the third gets the first newline block")
    run([[5, 32, 6]], 1, "full", -2, "strategy full keeps the class token")
    run([[5, 6, 7]], 2, "default", -2, "no <image> token: text only although pixel_values are given")
    run([[5, 32, 6]], 2, "other", -2, "unknown strategy -> ValueError")
    with open(OUT, "w") as f:
        json.dump(golden, f)
    print("wrote", OUT, len(golden["cases"]), "cases")


if __name__ == "__main__":
    main()
