#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "two_images or falls_back" > gpurun_out/t_new.log 2>&1
echo "tests exit $?"; tail -30 gpurun_out/t_new.log
