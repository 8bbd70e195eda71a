#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/tc_check.py tiny 4 > gpurun_out/tc_tiny.log 2>&1
echo "tiny exit $?"
tail -30 gpurun_out/tc_tiny.log
timeout -s KILL 300 python tools/tc_check.py wide2 6 > gpurun_out/tc_wide2.log 2>&1
echo "wide2 exit $?"
tail -30 gpurun_out/tc_wide2.log
