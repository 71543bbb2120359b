#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/c18_c2_launches.csv env N=200000 STEPS=1 python scripts/profile_run.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/c18_c3_launches.csv env N=100000 C3_CPU=0 python tests/probes/c3_probe.py > /dev/null 2>&1
