#!/bin/bash
O=gpurun_out/last; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "whole_read or exact_path or nodeless" > $O/gpu_tests_shortcut.log 2>&1; tail -n 2 $O/gpu_tests_shortcut.log
timeout 100 python bench.py --steps 5 --warmup 3 > $O/r2_bench_c2_n1.json 2> $O/bench.err; tail -c 300 $O/r2_bench_c2_n1.json; tail -n 3 $O/bench.err
