#!/bin/bash
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c7_*.log
for v in "" _gw16; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  echo "== variant '$v'" >> gpurun_out/c7_perf.log
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_parity.py -m gpu -x -q > gpurun_out/c7_tests$v.log 2>&1
  timeout 300 env N=200000 STEPS=3 python scripts/profile_run.py >> gpurun_out/c7_perf.log 2>&1
  timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py >> gpurun_out/c7_perf.log 2>&1
  timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
     --clock-control none -k regex:k_align -c 1 env N=100000 STEPS=1 python scripts/profile_run.py 2>&1 | grep -E "smsp__|gpu__time|dram__" >> gpurun_out/c7_perf.log
done
for f in gpurun_out/c7_tests*.log; do echo $f; tail -n 3 $f; done; cat gpurun_out/c7_perf.log
