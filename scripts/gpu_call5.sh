#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in "" _gw16; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_align -c 1 -f -o gpurun_out/r2_v1$v \
    env N=100000 STEPS=1 python scripts/profile_run.py > gpurun_out/ncu_v1$v.log 2>&1
  timeout 300 ncu -i gpurun_out/r2_v1$v.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_v1${v}_source.csv 2>/dev/null
  timeout 300 ncu -i gpurun_out/r2_v1$v.ncu-rep --page details > gpurun_out/r2_v1${v}_details.txt 2>/dev/null
  rm -f gpurun_out/r2_v1$v.ncu-rep
done
