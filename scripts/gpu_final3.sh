#!/bin/bash
# last build of the round: GPU tests, smoke(), the default bench line
O=gpurun_out/final3; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q > $O/r2_gpu_tests.log 2>&1; tail -n 2 $O/r2_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/smoke.log 2>&1; tail -n 1 $O/smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r2_bench_c2_n1.json 2> $O/bench.err; tail -c 400 $O/r2_bench_c2_n1.json
