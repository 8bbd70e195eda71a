#!/bin/bash
mkdir -p gpurun_out
for f in 1 2 3 8; do
B200_TC_INFLIGHT=$f timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mega-mode 2 > gpurun_out/bench_f$f.json 2> gpurun_out/bench_f$f.err
echo "bench inflight $f exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_f*.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
B200_TC_INFLIGHT=2 timeout -s KILL 300 python tools/mega_timeline.py 2 > gpurun_out/timeline_tc.txt 2>&1
head -24 gpurun_out/timeline_tc.txt | grep -v "producer issue\|consumer gate"
