#!/usr/bin/env python
"""One-screen summary of an `ncu -i X.ncu-rep --page raw --csv` export (first profiled kernel):
duration, DRAM bytes / throughput, tensor-pipe and TMEM activity, occupancy, launch geometry.
usage: python tools/ncu_summary.py gpurun_out/prof/*.raw.csv > profiles/r2_ncu_full_summary.txt"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
for path in sys.argv[1:]:
    with open(path) as f:
        rd = list(csv.reader(f))
    if len(rd) < 3:
        print(f"== {path}: empty")
        continue
    hdr, units, vals = rd[0], rd[1], rd[2]
    col = {h: i for i, h in enumerate(hdr)}
    name = vals[col["Kernel Name"]] if "Kernel Name" in col else "?"
    print(f"== {path.split('/')[-1]}  kernel: {name[:100]}")
    for k in KEYS:
        if k in col:
            print(f"   {k:88s} {vals[col[k]]:>16s} {units[col[k]]}")
    print()
