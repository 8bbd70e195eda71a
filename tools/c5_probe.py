#!/usr/bin/env python
"""Run only the lock-step legs of bench.py (C5: Qwen2-VL-7B, 8 rows; the 2B model with 8 rows) and print their numbers —
an A/B probe for decode_batch.cu changes (e.g. B200_BD_PREFETCH_MB=0 vs the default).
usage: python tools/c5_probe.py [--c5-out 256]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c5-rows", type=int, default=8)
    ap.add_argument("--c5-out", type=int, default=512)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = bench.c5_leg(1, 0, dev, args)
    print(json.dumps({"env": os.environ.get("B200_BD_PREFETCH_MB"), "c5": {k: out[k] for k in out if k != "workload"}}))


if __name__ == "__main__":
    main()
