#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do
B200_MEGA_FLOW=$f timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flow$f.json 2> gpurun_out/bench_flow$f.err
echo "bench flow $f exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_flow?.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
B200_MEGA_FLOW=1 timeout -s KILL 200 python tools/mega_timeline.py 1 2>&1 | head -12
