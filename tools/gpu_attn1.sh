#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention -x > gpurun_out/t_attn.log 2>&1
echo "attn tests exit $?"; tail -25 gpurun_out/t_attn.log
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_a.json')); print('decode tok/s', round(d['value'],1), 'prefill ms', round(d['prefill_ms'],2), 'img tok/s', round(d['prefill_img_tokens_per_sec'],0), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/bench_a.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:attention" -c 130 --csv --log-file gpurun_out/launches_attn.csv python tools/profile_decode.py 1 > gpurun_out/prof_attn.log 2>&1
python tools/agg_launches.py gpurun_out/launches_attn.csv | head -8
