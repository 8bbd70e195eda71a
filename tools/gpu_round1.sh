#!/bin/bash
# first GPU bring-up: op tests (non-GEMM, then GEMM), then engine tests; each under its own timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not gemm" -s > gpurun_out/t_ops.log 2>&1
echo "ops exit $?" >> gpurun_out/summary.txt
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm" -s > gpurun_out/t_gemm.log 2>&1
echo "gemm exit $?" >> gpurun_out/summary.txt
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -s -k "not full_size" > gpurun_out/t_engine.log 2>&1
echo "engine exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t_ops.log gpurun_out/t_gemm.log gpurun_out/t_engine.log
