#!/usr/bin/env python
"""Where does a slow `prefill_rows` call spend its host time?  Wraps the calls it makes with wall-clock timers.
usage: python tools/c3_step_probe2.py [iterations]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

ACC = {}


def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name

    def timed(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            ACC[label] = ACC.get(label, 0.0) + (time.perf_counter() - t) * 1e3
    setattr(obj, name, timed)


def main():
    from mlx_vlm_b200.models.llava import Model
    from mlx_vlm_b200.models.llava.config import llava_15_7b_config
    n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda", 0)
    cfg = llava_15_7b_config()
    model = Model(cfg, device=dev).init_random(2)
    eng, lm = model.engine, model.language_model
    wrap(lm, "_bind")
    wrap(eng, "prefill_batch")
    wrap(eng, "fetch_tokens")
    wrap(lm, "resolve_position_ids")
    wrap(eng.stream, "synchronize", "stream.synchronize")
    orig_pin = torch.Tensor.pin_memory

    def pin(self, *a, **k):
        t = time.perf_counter()
        try:
            return orig_pin(self, *a, **k)
        finally:
            ACC["pin_memory"] = ACC.get("pin_memory", 0.0) + (time.perf_counter() - t) * 1e3
    torch.Tensor.pin_memory = pin
    v = cfg.vision_config
    B, n_text = 8, 32
    P = (v.image_size // v.patch_size) ** 2
    rng = np.random.default_rng(11)
    pv_host = torch.from_numpy(rng.standard_normal((B, 3, v.image_size, v.image_size)).astype(np.float32)).pin_memory()
    text = rng.integers(3, 31000, size=n_text)
    ids = np.concatenate([text[:n_text // 2], np.full(P, cfg.image_token_index), text[n_text // 2:]])[None]
    T = ids.shape[1]
    for it in range(n_it):
        with torch.cuda.stream(eng.stream):
            pv = pv_host.to(dev, non_blocking=True)
        feats = model.encode_image(pv)
        embs = [model.get_input_embeddings(ids, pv, cached_image_features=feats[b:b + 1]).inputs_embeds for b in range(B)]
        rows, _ = lm.make_batch_cache(B, T + 8)
        caches = [lm.make_cache_row(rows.pool, b) for b in range(B)]
        torch.cuda.synchronize()
        ACC.clear()
        t = time.perf_counter()
        lm.prefill_rows([ids] * B, embs, caches, reserve_tokens=T + 8)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) * 1e3
        print(f"it {it:2d}: prefill_rows {dt:8.2f} ms  " + "  ".join(f"{k} {v:.2f}" for k, v in sorted(ACC.items())), flush=True)


if __name__ == "__main__":
    main()
