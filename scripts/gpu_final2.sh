#!/bin/bash
# final build: GPU tests, the bench line, the launch list of the bench command, kernel metrics at bench scale
O=gpurun_out/final2; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_gpu_tests.log 2>&1; tail -n 2 $O/r2_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_bench_c2_n1.json 2> $O/bench.err; tail -c 300 $O/r2_bench_c2_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
MGB_NO_EXACT_SHORTCUT=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
   --clock-control none -k regex:"k_align|k_seed|k_prepare" -c 3 --csv --log-file $O/r2_bench_scale_metrics.csv python scripts/profile_bench.py > $O/profile_bench.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_seed" -c 1 -f -o $O/k_seed python scripts/profile_bench.py > $O/ncu_k_seed.log 2>&1
ncu -i $O/k_seed.ncu-rep --page details > $O/r2_k_seed_details.txt 2>/dev/null; rm -f $O/k_seed.ncu-rep
ls $O
