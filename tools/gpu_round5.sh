#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full_size" > gpurun_out/t_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
for f in "" "--no-pdl" "--no-graph"; do
  timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $f > gpurun_out/bench$f.json 2>> gpurun_out/bench.err
  echo "bench $f exit $?" >> gpurun_out/summary.txt
done
K='regex:k_stream|k_attn|k_sample'
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 600 -c 300 --csv --log-file gpurun_out/launches_decode_r4.csv python tools/profile_decode.py 3 > gpurun_out/prof.log 2>&1
cat gpurun_out/summary.txt
grep -E "passed|failed" gpurun_out/t_engine.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*.json')):
    try:
        d=json.load(open(f)); print(f, 'decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), d['e2e']['generation_tps_api'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/bench.err
python tools/agg_launches.py gpurun_out/launches_decode_r4.csv
