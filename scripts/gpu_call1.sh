#!/bin/bash
# round-2 GPU call 1: box facts, source-level ncu captures of k_align (C2 and C3 shapes), probes of what round 1 left unmeasured
set -x
mkdir -p gpurun_out
{ nvidia-smi; free -g; nproc; lscpu | head -20; } > gpurun_out/box.txt 2>&1
export PYTHONUNBUFFERED=1
# C2 shape, small: source-level profile of k_align
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_align -c 1 -f -o gpurun_out/r2_base_c2 \
    env N=100000 STEPS=1 python scripts/profile_run.py > gpurun_out/ncu_c2.log 2>&1
timeout 300 ncu -i gpurun_out/r2_base_c2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_base_c2_source.csv 2>/dev/null
timeout 300 ncu -i gpurun_out/r2_base_c2.ncu-rep --page details > gpurun_out/r2_base_c2_details.txt 2>/dev/null
rm -f gpurun_out/r2_base_c2.ncu-rep
# C3 shape: the same for the sub-k / error path
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_align -c 1 -f -o gpurun_out/r2_base_c3 \
    env N=50000 C3_CPU=0 python tests/probes/c3_probe.py > gpurun_out/ncu_c3.log 2>&1
timeout 300 ncu -i gpurun_out/r2_base_c3.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2_base_c3_source.csv 2>/dev/null
timeout 300 ncu -i gpurun_out/r2_base_c3.ncu-rep --page details > gpurun_out/r2_base_c3_details.txt 2>/dev/null
rm -f gpurun_out/r2_base_c3.ncu-rep
# probes
timeout 600 env N=200000 C3_CPU_READS=8000 python tests/probes/c3_probe.py > gpurun_out/c3_probe.log 2>&1
timeout 900 python tests/probes/modes_probe.py > gpurun_out/modes_probe.log 2>&1
timeout 900 env N=1000000 python tests/probes/c4_probe.py > gpurun_out/c4_probe.log 2>&1
ls -la gpurun_out
