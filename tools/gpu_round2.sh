#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout -s KILL 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full_size" > gpurun_out/t_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.json 2>> gpurun_out/bench.err
cat gpurun_out/summary.txt
grep -E "rel_l2=|\|cuda|passed|failed|tokens" gpurun_out/t_engine.log | cut -c1-220
cat gpurun_out/bench.json; cat gpurun_out/bench_nograph.json; tail -n 20 gpurun_out/bench.err
