#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/mega_timeline.py 1 > gpurun_out/mega_timeline_r1.txt 2>&1
head -8 gpurun_out/mega_timeline_r1.txt
timeout -s KILL 200 python tools/mega_timeline.py 4 > gpurun_out/mega_timeline_flow_r1.txt 2>&1
head -5 gpurun_out/mega_timeline_flow_r1.txt
