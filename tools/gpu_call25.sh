#!/bin/bash
mkdir -p gpurun_out
echo "== base"; timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c25.jsonl --configs "mode=6,warps_per_cta=8;mode=6,warps_per_cta=10" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== p4"; PR_LIB_PATH=$PWD/build_variants/lib_p4.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_p4.jsonl --configs "mode=6,warps_per_cta=10;mode=6,warps_per_cta=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
