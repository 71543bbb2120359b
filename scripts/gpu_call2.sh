#!/bin/bash
# round-2 GPU call 2: new extender (single-loop, compact table format, generic lane-group width): parity + timing
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests_gw32.log 2>&1
for v in "" _gw16 _gw8; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  echo "== variant '$v'" >> gpurun_out/c2_perf.log
  if [ -n "$v" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_parity.py -m gpu -x -q > gpurun_out/c2_tests$v.log 2>&1; fi
  timeout 300 env N=200000 STEPS=3 python scripts/profile_run.py >> gpurun_out/c2_perf.log 2>&1
  timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py >> gpurun_out/c2_perf.log 2>&1
done
tail -3 gpurun_out/c2_tests*.log; cat gpurun_out/c2_perf.log
