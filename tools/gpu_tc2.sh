#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/tc_check2.py tiny1 2 > gpurun_out/tc2_tiny.log 2>&1
echo "tiny1 exit $?"; tail -30 gpurun_out/tc2_tiny.log
timeout -s KILL 300 python tools/tc_check2.py wide1 8 > gpurun_out/tc2_wide.log 2>&1
echo "wide1 exit $?"; tail -30 gpurun_out/tc2_wide.log
