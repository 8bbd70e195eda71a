#!/usr/bin/env python
"""Per-call host wall + device times of the C3 step over many iterations (looks for drift across iterations).
usage: python tools/c3_step_probe.py [iterations]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from mlx_vlm_b200.models.llava import Model
    from mlx_vlm_b200.models.llava.config import llava_15_7b_config
    n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dev = torch.device("cuda", 0)
    cfg = llava_15_7b_config()
    model = Model(cfg, device=dev).init_random(2)
    eng, lm = model.engine, model.language_model
    v = cfg.vision_config
    B, n_text = 8, 32
    P = (v.image_size // v.patch_size) ** 2
    rng = np.random.default_rng(11)
    pv_host = torch.from_numpy(rng.standard_normal((B, 3, v.image_size, v.image_size)).astype(np.float32)).pin_memory()
    text = rng.integers(3, 31000, size=n_text)
    ids = np.concatenate([text[:n_text // 2], np.full(P, cfg.image_token_index), text[n_text // 2:]])[None]
    T = ids.shape[1]

    def timed(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, (time.perf_counter() - t) * 1e3

    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    except Exception:
        nv = None
    t_start = time.perf_counter()
    for it in range(n_it):
        with torch.cuda.stream(eng.stream):
            pv = pv_host.to(dev, non_blocking=True)
        feats, t_enc = timed(lambda: model.encode_image(pv))
        embs, t_emb = timed(lambda: [model.get_input_embeddings(ids, pv, cached_image_features=feats[b:b + 1]).inputs_embeds
                                     for b in range(B)])
        (rows, _), t_cache = timed(lambda: lm.make_batch_cache(B, T + 8))
        caches, t_rows = timed(lambda: [lm.make_cache_row(rows.pool, b) for b in range(B)])
        _, t_pre = timed(lambda: lm.prefill_rows([ids] * B, embs, caches, reserve_tokens=T + 8))
        free, total = torch.cuda.mem_get_info()
        hw = ""
        if nv is not None:
            hw = (f"  t+{time.perf_counter() - t_start:5.2f}s  {pynvml.nvmlDeviceGetPowerUsage(nv) / 1000:5.0f} W  "
                  f"sm {pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM)} MHz  "
                  f"reasons 0x{pynvml.nvmlDeviceGetCurrentClocksEventReasons(nv):x}  {pynvml.nvmlDeviceGetTemperature(nv, 0)} C")
        print(f"it {it:2d}: encode {t_enc:7.2f}  embed x8 {t_emb:7.2f}  make_batch_cache {t_cache:7.2f}  rows {t_rows:6.2f}  "
              f"prefill_rows {t_pre:7.2f} ms   torch alloc {torch.cuda.memory_allocated() / 2**30:6.2f} GiB reserved "
              f"{torch.cuda.memory_reserved() / 2**30:6.2f} GiB  device free {free / 2**30:6.1f} GiB" + hw, flush=True)


if __name__ == "__main__":
    main()
