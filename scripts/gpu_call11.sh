#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c11_tests.log 2>&1; tail -n 3 gpurun_out/c11_tests.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_n1_v1.json 2> gpurun_out/c11_bench_c2.err; tail -c 3000 gpurun_out/r2_bench_c2_n1_v1.json; tail -n 5 gpurun_out/c11_bench_c2.err
BENCH_CONFIG=c3 BENCH_GENOME=20000000 BENCH_READS=400000 timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/c11_bench_c3_small.json 2> gpurun_out/c11_bench_c3.err; tail -c 2500 gpurun_out/c11_bench_c3_small.json; tail -n 5 gpurun_out/c11_bench_c3.err
