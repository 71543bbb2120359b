#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_parity.py tests/test_primary_gpu.py -m gpu -x -q 2>&1 | tail -n 2
timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py 2>&1 | tail -n 1
timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py 2>&1 | tail -n 1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none -k regex:"k_seed" -c 1 python scripts/profile_bench.py 2>&1 | grep -E "__"
