#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py 2>&1 | tail -n 1
timeout 300 env N=100000 C3_CPU=0 python tests/probes/c3_probe.py 2>&1 | tail -n 1
FUZZ_SECONDS=40 timeout 300 python tests/probes/fuzz_sweep.py 80000 81000 2>&1 | tail -n 1
