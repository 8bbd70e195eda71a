#!/bin/bash
# Launch lists of the PRODUCTION configurations: a plain run measures the GEMM configurations and writes them to a
# tuning file, the ncu runs read it back (no measurement under the profiler); --cache-control none keeps the L2 warm
# like in the real sequence (kernels are still serialised: no PDL overlap).
set -u
mkdir -p gpurun_out/prof
P=gpurun_out/prof
export B200_WT_TUNE_FILE=$PWD/$P/wt_tune.txt
rm -f $B200_WT_TUNE_FILE
python tools/profile_decode.py 3 > $P/plain_c2.out 2>&1
python tools/profile_batch.py 4 8 500 4 > $P/plain_batch.out 2>&1
wc -l $B200_WT_TUNE_FILE
NCU="ncu --clock-control none --cache-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_c2_warm.csv -c 8000 python tools/profile_decode.py 3 > $P/launches_c2_warm.out 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_batch_warm.csv -c 3000 python tools/profile_batch.py 4 8 500 4 > $P/launches_batch_warm.out 2>&1
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $P/launches_batch_cold.csv -c 3000 python tools/profile_batch.py 4 8 500 4 > $P/launches_batch_cold.out 2>&1
