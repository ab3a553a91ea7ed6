#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:bm25_lean -s 100 -c 3 --csv --log-file gpurun_out/l2_c38.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/l2_c38.log 2>&1; echo "rc=$?"
grep -v "^==" gpurun_out/l2_c38.csv | cut -d, -f5,13- | tail -24
