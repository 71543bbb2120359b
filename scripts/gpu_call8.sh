#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c8_*.log
for v in _mb3 "" _mb5 _mb6; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  echo "== variant '$v'" >> gpurun_out/c8_perf.log
  timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py >> gpurun_out/c8_perf.log 2>&1
  timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py >> gpurun_out/c8_perf.log 2>&1
done
cat gpurun_out/c8_perf.log
