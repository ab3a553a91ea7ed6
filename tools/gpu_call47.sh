#!/bin/bash
mkdir -p gpurun_out
for v in s12p3 s12p4 s12p6; do echo "== sweep $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c47_$v.jsonl --configs "mode=8,warps_per_cta=4,subs_per_item=12;mode=8,warps_per_cta=4,subs_per_item=24;mode=8,warps_per_cta=12,subs_per_item=12" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300; done
