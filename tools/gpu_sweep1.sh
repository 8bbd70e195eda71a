#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sw_*.json
run() { # name mode envs...
  name=$1; mode=$2; shift 2
  env "$@" timeout -s KILL 150 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --mega-mode $mode > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err
  echo "$name exit $?"
}
run fma_base 1 X=1
run fma_kv 1 B200_KV_KEEP=1
run fma_if2 1 B200_FMA_INFLIGHT=2
run fma_kv_if2 1 B200_KV_KEEP=1 B200_FMA_INFLIGHT=2
run tc_base 2 X=1
run tc_kv 2 B200_KV_KEEP=1
run tc_kv_pf1 2 B200_KV_KEEP=1 B200_L2_PREFETCH=1
run tc_kv_pf2 2 B200_KV_KEEP=1 B200_L2_PREFETCH=2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sw_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], 'tok/s', round(d['value'],1), 'ms/tok', round(d['decode_ms_per_token'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
    except Exception as e: print(f, 'ERR', e)
PY
