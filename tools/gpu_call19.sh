#!/bin/bash
# Sub-tile size variants (PR_SUB_SHIFT 10/12) of modes 4/6; mode 7 with costlier rescoring.
mkdir -p gpurun_out
echo "== sub12"; PR_LIB_PATH=$PWD/build_variants/lib_sub12.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_sub12.jsonl --configs "mode=6,warps_per_cta=4,subs_per_item=6;mode=6,warps_per_cta=8,subs_per_item=6;mode=4,warps_per_cta=4,subs_per_item=6;mode=4,warps_per_cta=8,subs_per_item=6;mode=4,warps_per_cta=4,subs_per_item=6,lazy_zero=1;mode=6,warps_per_cta=4,subs_per_item=12" 2>&1 | grep -v "^\[bench" | cut -c1-300
echo "== sub12p4"; PR_LIB_PATH=$PWD/build_variants/lib_sub12p4.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_sub12p4.jsonl --configs "mode=6,warps_per_cta=4,subs_per_item=6;mode=6,warps_per_cta=8,subs_per_item=6" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== sub10"; PR_LIB_PATH=$PWD/build_variants/lib_sub10.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_sub10.jsonl --configs "mode=6,warps_per_cta=8,subs_per_item=24;mode=4,warps_per_cta=16,subs_per_item=24;mode=4,warps_per_cta=8,subs_per_item=24" 2>&1 | grep -v "^\[bench" | cut -c1-300
echo "== mode7 rescoring cost"; timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c19.jsonl --configs "mode=7,rescore_cost=1024;mode=7,rescore_cost=8192" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
