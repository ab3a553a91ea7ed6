#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== pytest gpu (all but full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c40.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_c40.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c40.jsonl --configs "mode=8;mode=8,warps_per_cta=10;mode=8,warps_per_cta=12;mode=8,subs_per_item=12;mode=8,subs_per_item=48;mode=8,docs_per_launch=786432;mode=8,docs_per_launch=786432,subs_per_item=48;mode=8,docs_per_launch=1572864,subs_per_item=32" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-400
echo "== ncu full"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 30 -c 1 -o gpurun_out/prof_lean_c40 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lean_c40.log 2>&1; echo "rc=$?"
