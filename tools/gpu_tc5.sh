#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mega-mode 2 > gpurun_out/bench_m2.json 2> gpurun_out/bench_m2.err
echo "bench exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_m2.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
timeout -s KILL 200 python tools/mega_timeline.py 2 > gpurun_out/timeline_tc.txt 2>&1
grep -v "producer issue\|consumer gate" gpurun_out/timeline_tc.txt | head -30
timeout -s KILL 200 python tools/tc_check.py wide2 4 | tail -4
