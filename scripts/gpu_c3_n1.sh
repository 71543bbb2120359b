#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
free -g | head -2
BENCH_CONFIG=c3 timeout 2400 python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; tail -c 3000 gpurun_out/r2_bench_c3_n1.json; tail -n 8 gpurun_out/r2_bench_c3_n1.err
