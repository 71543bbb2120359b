#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in "" _mb1; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  timeout 600 ncu --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section InstructionStats --section MemoryWorkloadAnalysis_Tables --section MemoryWorkloadAnalysis \
     --metrics smsp__inst_executed_pipe_uniform.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed_op_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_active.avg,idc__requests.sum,idc__requests_lookup_miss.sum \
     --clock-control none -k regex:k_align -c 1 env N=100000 STEPS=1 python scripts/profile_run.py > gpurun_out/c14_ncu$v.txt 2>&1
done
