"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        rows.append((x["Kernel Name"], v, x.get("Grid Size"), x.get("Block Size")))
agg = collections.OrderedDict()
for k, v, g, b in rows:
    key = (re.sub(r"\(.*", "", k)[:48], g)
    a = agg.setdefault(key, [0, 0.0, 1e9, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(v for _, v, _, _ in rows)
print(f"{len(rows)} launches, total {tot:.1f} us")
for k, (n, t, lo, hi) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:50s} grid={k[1]:16s} n={n:5d} total={t:9.1f}us avg={t/n:8.2f} min={lo:7.2f} max={hi:8.2f} share={100*t/tot:5.1f}%")
