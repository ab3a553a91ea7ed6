#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv
echo "== new"; timeout 900 python tools/sweep.py --reps 3 --out gpurun_out/sweep_c50.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== old"; PR_LIB_PATH=$PWD/build_variants/lib_old.so timeout 900 python tools/sweep.py --reps 3 --out gpurun_out/sweep_c50_old.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== new again"; timeout 900 python tools/sweep.py --reps 3 --out gpurun_out/sweep_c50.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
