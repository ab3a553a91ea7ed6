#!/bin/bash
mkdir -p gpurun_out
for v in el1 el2; do echo "== sweep $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c39_$v.jsonl --configs "mode=8;mode=8,docs_per_launch=393216;mode=8,docs_per_launch=196608,subs_per_item=48" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300; done
echo "== L2 metrics el2"; PR_LIB_PATH=$PWD/build_variants/lib_el2.so timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none -k regex:bm25_lean -s 100 -c 1 --csv --log-file gpurun_out/l2_c39.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/l2_c39.log 2>&1; grep -v "^==" gpurun_out/l2_c39.csv | cut -d, -f13- | tail -3
