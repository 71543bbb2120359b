#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c20_tests.log 2>&1; tail -n 3 gpurun_out/c20_tests.log
timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py 2>&1 | tail -n 1
timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py 2>&1 | tail -n 1
