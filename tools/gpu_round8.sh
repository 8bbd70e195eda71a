#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/bench*.json gpurun_out/summary.txt gpurun_out/bench.err
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full_size" > gpurun_out/t_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
for f in "" "--no-graph" "--no-mega"; do
  timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $f > gpurun_out/bench$f.json 2>> gpurun_out/bench.err
  echo "bench $f exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
grep -E "passed|failed|Error|error" gpurun_out/t_engine.log | head
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench.err
