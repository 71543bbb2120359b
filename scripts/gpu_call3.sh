#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in "" _gw16; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
     --clock-control none -k regex:k_align -c 1 env N=100000 STEPS=1 python scripts/profile_run.py > gpurun_out/c3_div$v.log 2>&1
done
cat gpurun_out/c3_div*.log | grep -E "smsp__|gpu__time|sm__warps|dram__"
