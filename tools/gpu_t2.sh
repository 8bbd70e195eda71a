#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "not full_size" > gpurun_out/t_new.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/t_new.log
for m in 1 2; do
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --mega-mode $m > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err
echo "bench mode $m exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_m?.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
timeout -s KILL 200 python tools/mega_timeline.py 1 2>&1 | head -8
