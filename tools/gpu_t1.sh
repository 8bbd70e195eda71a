#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "batch_generator or tensor_core or generate_path or multi_kernel" > gpurun_out/t_new.log 2>&1
echo "exit $?"; tail -30 gpurun_out/t_new.log
