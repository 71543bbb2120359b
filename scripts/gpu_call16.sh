#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c16_perf.log
for v in "" _noinl; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  echo "== variant '$v'" >> gpurun_out/c16_perf.log
  MGB_NO_EXACT_SHORTCUT=1 timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py >> gpurun_out/c16_perf.log 2>&1
  timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py >> gpurun_out/c16_perf.log 2>&1
done
unset MGB_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_parity.py -m gpu -x -q 2>&1 | tail -n 2
cat gpurun_out/c16_perf.log
