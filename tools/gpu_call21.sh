#!/bin/bash
mkdir -p gpurun_out
for v in rem16 rem0 rem32h4 rem32h16; do
echo "== $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_$v.jsonl --configs "mode=6,warps_per_cta=8,subs_per_item=24;mode=6,warps_per_cta=8,subs_per_item=48,docs_per_launch=196608" 2>&1 | grep -v "^\[bench" | cut -c1-300
done
