#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x > gpurun_out/t_k.log 2>&1
echo "kernel tests exit $?"; tail -4 gpurun_out/t_k.log
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "not full_size" > gpurun_out/t_e.log 2>&1
echo "engine tests exit $?"; tail -4 gpurun_out/t_e.log
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_g.json')); print('decode tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'prefill ms', round(d['prefill_ms'],2), 'img tok/s', round(d['prefill_img_tokens_per_sec']), 'e2e', round(d['e2e']['value'],1))
PY
