#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c15_tests.log 2>&1; tail -n 3 gpurun_out/c15_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_n1_v2.json 2> gpurun_out/c15_bench_c2.err; tail -c 3500 gpurun_out/r2_bench_c2_n1_v2.json; tail -n 5 gpurun_out/c15_bench_c2.err
FUZZ_SECONDS=100 timeout 300 python tests/probes/fuzz_sweep.py 60000 61000 > gpurun_out/c15_fuzz.log 2>&1; tail -n 2 gpurun_out/c15_fuzz.log
