#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2"
BENCH_GENOME=20000000 BENCH_READS=300000 timeout 600 $RUN --steps 2 --warmup 1 > gpurun_out/c12_c2_n2.json 2> gpurun_out/c12_c2.err; tail -c 1800 gpurun_out/c12_c2_n2.json; tail -n 5 gpurun_out/c12_c2.err
BENCH_CONFIG=c3 BENCH_GENOME=20000000 BENCH_READS=400000 timeout 600 $RUN --steps 2 --warmup 1 > gpurun_out/c12_c3_n2.json 2> gpurun_out/c12_c3.err; tail -c 1800 gpurun_out/c12_c3_n2.json; tail -n 5 gpurun_out/c12_c3.err
