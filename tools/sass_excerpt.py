#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the shipped library uses
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy,
LDTM = tcgen05.ld from TMEM, SYNCS.* = mbarrier, UTCBAR = tcgen05.commit, FFMA2 = packed fp32 FMA).
usage: python tools/sass_excerpt.py [path/to/libb200vlm.so] > profiles/r2_sass_excerpts.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "mlx_vlm_b200", "libb200vlm.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCCP|SYNCS|UTMAPF|ACQBULK|"
                 r"FFMA2|MUFU\.EX2|HMMA|UTCATOMSWS|REDG|ATOMG)[\.\w]*")
counts, first, fn = collections.defaultdict(collections.Counter), {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = pat.search(line)
    if m and fn:
        counts[fn][m.group(0)] += 1
        first.setdefault((fn, m.group(1)), line.strip())
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {os.path.basename(so)} : mnemonic counts per kernel (sm_100a)\n")
tot = collections.Counter()
for fn, nm in sorted(zip(counts, names), key=lambda kv: -sum(counts[kv[0]].values())):
    c = counts[fn]
    tot.update(c)
    print(nm[:160])
    print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
    for key in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM"):
        if (fn, key) in first:
            print("      e.g. " + first[(fn, key)][:150])
print("\n# totals: " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
