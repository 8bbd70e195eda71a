#!/bin/bash
# Round-2 profiling run on ONE B200 (gpurun): launch lists of the final code + `--set full` captures of the
# top kernels.  Everything lands in gpurun_out/prof/; tools/agg_launches.py + the summaries in profiles/ are
# made from it here.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out/prof
P=gpurun_out/prof
NCU="ncu --clock-control none --cache-control none"   # warm L2: durations comparable with the CUDA-event numbers
# launch lists (gpu__time_duration per launch; cold-cache, serialised -> shares only)
$NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_c2.csv -c 8000 python tools/profile_decode.py 3 > $P/launches_c2.out 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $P/launches_batch.csv -c 3000 python tools/profile_batch.py 4 8 500 4 > $P/launches_batch.out 2>&1
# full captures
full() {  # name, kernel regex, skip, target...
  local name=$1 k=$2 s=$3; shift 3
  timeout -s KILL 300 $NCU --set full --import-source on -k regex:$k -s $s -c 1 -f -o $P/$name "$@" > $P/$name.out 2>&1
  ncu -i $P/$name.ncu-rep --page raw --csv > $P/$name.raw.csv 2>/dev/null
}
full k_mega k_mega 2 python tools/profile_decode.py 3
full gemm_wt_vit_fc1 gemm_wt_kernel 2 python tools/profile_ops.py vit_fc1
full gemm_wt_lm_gateup gemm_wt_kernel 2 python tools/profile_ops.py lm_gateup
full gemm_wt_dec8_gateup gemm_wt_kernel 2 python tools/profile_ops.py dec8_gateup
full gemm_wt_clip_fc1x8 gemm_wt_kernel 2 python tools/profile_ops.py clip_fc1x8
full attention_fa_vit attention_fa_kernel 2 python tools/profile_ops.py fa_vit
full bd_attn bd_attn_kernel 6 python tools/profile_batch.py 4 8 500 4
full attention_f32_clip attention_f32_kernel 2 python tools/profile_llava.py 2 8 2
full finish_rows finish_rows_kernel 4 python tools/profile_batch.py 4 8 500 4
ls -la $P | head -40
