#!/bin/bash
# ncu --set full of one full-size (393,216-document) launch of the lean kernel in the middle of a bench step
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s ${1:-85} -c 1 -o gpurun_out/prof_lean_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lean_final.log 2>&1; echo "rc=$?"
