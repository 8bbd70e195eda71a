#!/usr/bin/env python
"""Per-kernel totals of the LAST request / step in an ncu launch list (gpu__time_duration.sum, csv):
   python tools/last_request.py launches.csv c2     -> vision tower, LM prefill, decode steps of the last C2 request
   python tools/last_request.py launches.csv batch  -> the last lock-step decode step, launch by launch"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", x["Kernel Name"]).split("::")[-1]
        name = re.sub(r"<.*", "", name)[:34]
        rows.append((name, float(x["Metric Value"].replace(",", "")) / 1e3, x["Grid Size"]))
kind = sys.argv[2] if len(sys.argv) > 2 else "c2"


def table(part, title):
    agg = collections.OrderedDict()
    for n, t, g in part:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"{title}: {len(part)} launches, {tot:.1f} us")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {n:34s} x{c:4d} {t:9.1f} us  avg {t / c:7.2f}  {100 * t / tot:5.1f} %")


names = [r[0] for r in rows]
if kind == "batch":
    idx = [i for i, n in enumerate(names) if n.startswith("bd_sample")]
    a, b = idx[-2] + 1, idx[-1] + 1
    for n, t, g in rows[a:b]:
        print(f"{n:34s} {g:16s} {t:8.2f}")
    table(rows[a:b], "last lock-step step")
else:
    merges = [i for i, n in enumerate(names) if n == "embed_merge_kernel"]
    last = merges[-1]
    j = last - 1
    while j >= 0 and names[j] not in ("k_sample", "k_mega", "embed_merge_kernel"):
        j -= 1
    table(rows[j + 1:last], "vision tower (last request)")
    k = last
    while k < len(names) and names[k] != "k_sample":
        k += 1
    table(rows[last:k + 1], "merge + LM prefill (last request)")
    table([r for r in rows[k + 1:] if r[0] == "k_mega"], "decode steps after it")
