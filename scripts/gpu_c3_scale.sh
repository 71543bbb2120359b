#!/bin/bash
# usage: gpu_c3_scale.sh N  -- BENCH_CONFIG=c3 at N GPUs (strong scaling, 10 M reads in total)
N=$1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
BENCH_CONFIG=c3 timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err
tail -c 2500 gpurun_out/r2_bench_c3_n$N.json; tail -n 6 gpurun_out/r2_bench_c3_n$N.err
