#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c13_*.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c13_tests.log 2>&1; tail -n 3 gpurun_out/c13_tests.log
timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py >> gpurun_out/c13_perf.log 2>&1
timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py >> gpurun_out/c13_perf.log 2>&1
timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_requests.sum,sm__icc_requests_lookup_miss.sum,dram__bytes_read.sum,dram__bytes_write.sum \
     --clock-control none -k regex:k_align -c 1 env N=100000 STEPS=1 python scripts/profile_run.py 2>&1 | grep -E "__" >> gpurun_out/c13_perf.log
cat gpurun_out/c13_perf.log
