#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full_size" > gpurun_out/t_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
# launch list: second iteration only (skip the first iteration's launches: 2+32*8+3 vit, 28*9+3 prefill, 3*142 decode)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 950 -c 950 --csv --log-file gpurun_out/launches_r1.csv python tools/profile_decode.py 3 > gpurun_out/prof.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
grep -E "rel_l2=|\|cuda|passed|failed|tokens" gpurun_out/t_engine.log | cut -c1-200
tail -3 gpurun_out/prof.log
