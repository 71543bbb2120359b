#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for g in 32 0; do
  echo "== MGB_L2_FETCH=$g"
  MGB_L2_FETCH=$g MGB_NO_EXACT_SHORTCUT=1 timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py 2>&1 | tail -n 1
  MGB_L2_FETCH=$g timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py 2>&1 | tail -n 1
done
MGB_L2_FETCH=32 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_seed|k_align" -c 2 python scripts/profile_bench.py 2>&1 | grep -E "__|k_"
