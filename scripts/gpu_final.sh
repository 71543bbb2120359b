#!/bin/bash
# round-2 final evidence: GPU tests, the bench line, the ncu launch list of the bench command, --set full captures of the
# dominant kernels, DRAM traffic at bench scale, probes of the other configs. Outputs under gpurun_out/final/.
O=gpurun_out/final; mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > $O/r2_gpu_tests.log 2>&1; tail -n 2 $O/r2_gpu_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/r2_bench_c2_n1.json 2> $O/bench.err; tail -c 600 $O/r2_bench_c2_n1.json
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1
# DRAM traffic / instructions of the kernels at bench scale (1 M reads, G = 100 Mbp), shortcut off for k_align
MGB_NO_EXACT_SHORTCUT=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
   --clock-control none -k regex:"k_align|k_seed|k_premap|k_prepare" -c 8 --csv --log-file $O/r2_bench_scale_metrics.csv python scripts/profile_bench.py > $O/profile_bench.log 2>&1
# --set full, source-level, of k_align (C2 shape, shortcut off) and of k_seed / k_subk (C3 shape)
MGB_NO_EXACT_SHORTCUT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_align -c 1 -f -o $O/k_align env N=100000 STEPS=1 python scripts/profile_run.py > $O/ncu_k_align.log 2>&1
ncu -i $O/k_align.ncu-rep --page source --csv --print-source cuda,sass > $O/r2_k_align_source.csv 2>/dev/null
ncu -i $O/k_align.ncu-rep --page details > $O/r2_k_align_details.txt 2>/dev/null; rm -f $O/k_align.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:"k_seed|k_subk" -c 2 -f -o $O/k_seed env N=100000 C3_CPU=0 python tests/probes/c3_probe.py > $O/ncu_k_seed.log 2>&1
ncu -i $O/k_seed.ncu-rep --page details > $O/r2_k_seed_subk_details_c3.txt 2>/dev/null; rm -f $O/k_seed.ncu-rep
# other configs
{ echo "== c4 (protein, 1 M x 100 aa vs 50 M nodes)"; timeout 900 env N=1000000 python tests/probes/c4_probe.py; echo "== graph modes (C2 shape, 200 k reads, G = 20 Mbp)"; timeout 900 python tests/probes/modes_probe.py; } > $O/r2_probes.txt 2>&1
timeout 600 env COUNTS=1000,100000,1000000 python scripts/c5_microbench.py > $O/c5.log 2>&1; cp gpurun_out/c5_microbench.json $O/r2_c5_dp_microbench.json
ls -la $O
