#!/bin/bash
# short end-of-session check: all GPU tests, smoke, one bench line (no ncu)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu > gpurun_out/t_gpu_all.log 2>&1
echo "pytest gpu exit $?"; grep -E "passed|failed" gpurun_out/t_gpu_all.log | tail -2
timeout -s KILL 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json')); print('decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
