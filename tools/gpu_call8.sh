#!/bin/bash
mkdir -p gpurun_out
for v in s12_pf4 s12_pf2; do
echo "== $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_$v.jsonl --configs "mode=4,warps_per_cta=4,subs_per_item=6;mode=4,warps_per_cta=13,subs_per_item=6;mode=4,warps_per_cta=4,subs_per_item=6,lazy_zero=2" 2>&1 | grep -v bench | cut -c1-200
done
echo "== s11_pf2"; PR_LIB_PATH=$PWD/build_variants/lib_s11_pf2.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_s11_pf2.jsonl --configs "mode=4,warps_per_cta=8;mode=4,warps_per_cta=8,lazy_zero=2;mode=4,warps_per_cta=9" 2>&1 | grep -v bench | cut -c1-200
