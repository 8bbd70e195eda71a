#!/bin/bash
mkdir -p gpurun_out
for m in 2; do
timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mega-mode $m > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err
echo "bench mode $m exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_m*.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_m2.err
timeout -s KILL 300 python tools/mega_timeline.py 2 > gpurun_out/timeline_tc.txt 2>&1
head -60 gpurun_out/timeline_tc.txt; timeout -s KILL 300 python tools/tc_check.py wide2 4 | tail -8; timeout -s KILL 300 python tools/tc_check.py tiny 4 | tail -8
