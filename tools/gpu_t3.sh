#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "not full_size" > gpurun_out/t_new.log 2>&1
echo "tests exit $?"; tail -6 gpurun_out/t_new.log
for f in 1 0; do
B200_MEGA_FLOW=$f timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flow$f.json 2> gpurun_out/bench_flow$f.err
echo "bench flow $f exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_flow?.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_flow1.err
timeout -s KILL 200 python tools/mega_timeline.py 1 2>&1 | head -7
