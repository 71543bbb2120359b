#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/c9_*.log
for v in _mb1 _mb2; do
  export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb$v.so
  echo "== variant '$v'" >> gpurun_out/c9_perf.log
  timeout 300 env N=200000 STEPS=2 python scripts/profile_run.py >> gpurun_out/c9_perf.log 2>&1
done
export MGB_LIB=$PWD/metagraph_b200/_lib/libmgb.so
timeout 600 ncu --metrics sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests_lookup_miss_tag_miss.sum,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction_lookup_miss.sum,smsp__inst_executed.sum,gpu__time_duration.sum,sm__cycles_active.avg,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum \
     --clock-control none -k regex:k_align -c 1 env N=100000 STEPS=1 python scripts/profile_run.py 2>&1 | grep -E "__" >> gpurun_out/c9_perf.log
cat gpurun_out/c9_perf.log
