#!/usr/bin/env python
"""One warm batched prefill of the C3 request (8 x T=608 on the Llama-7B geometry) inside a cudaProfilerStart/Stop range,
for `ncu --profile-from-start off --metrics gpu__time_duration.sum` (launch list of the final code).
Prints the per-kernel totals itself when given an ncu csv:  python tools/c3_launch_list.py --summarise launches.csv"""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", x["Kernel Name"]).split("::")[-1])[:34]
        key = (name, x["Grid Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"].replace(",", "")) / 1e3
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"C3 batched LM prefill (8 x T=608, Llama-7B geometry), one call: {n} launches, {tot / 1e3:.2f} ms serialised under ncu")
    for (name, grid), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {name:34s} grid {grid:18s} x{c:4d} {t:10.1f} us  avg {t / c:8.2f}  {100 * t / tot:5.1f} %")


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        return summarise(sys.argv[2])
    import numpy as np
    import torch
    from mlx_vlm_b200.models.llava import Model
    from mlx_vlm_b200.models.llava.config import llava_15_7b_config
    dev = torch.device("cuda", 0)
    cfg = llava_15_7b_config()
    model = Model(cfg, device=dev).init_random(2)
    eng, lm = model.engine, model.language_model
    v = cfg.vision_config
    B, n_text = 8, 32
    P = (v.image_size // v.patch_size) ** 2
    rng = np.random.default_rng(11)
    pv = torch.from_numpy(rng.standard_normal((B, 3, v.image_size, v.image_size)).astype(np.float32)).to(dev)
    text = rng.integers(3, 31000, size=n_text)
    ids = np.concatenate([text[:n_text // 2], np.full(P, cfg.image_token_index), text[n_text // 2:]])[None]
    T = ids.shape[1]
    feats = model.encode_image(pv)
    embs = [model.get_input_embeddings(ids, pv, cached_image_features=feats[b:b + 1]).inputs_embeds for b in range(B)]
    for it in range(3):
        rows, _ = lm.make_batch_cache(B, T + 8)
        caches = [lm.make_cache_row(rows.pool, b) for b in range(B)]
        torch.cuda.synchronize()
        if it == 2:
            torch.cuda.profiler.start()
        lm.prefill_rows([ids] * B, embs, caches, reserve_tokens=T + 8)
        torch.cuda.synchronize()
        if it == 2:
            torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
