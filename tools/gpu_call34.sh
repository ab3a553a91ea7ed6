#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 100 -c 1 -o gpurun_out/prof_lean_c34 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lean_c34.log 2>&1; echo "rc=$?"
