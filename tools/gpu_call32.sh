#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu bm25 + prober"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_prober.py -m gpu -x -q > gpurun_out/pytest_gpu_c32.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu_c32.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c32.jsonl --configs "mode=8;mode=6;mode=8,warps_per_cta=10;mode=8,warps_per_cta=12" 2>&1 | grep -v "^\[bench" | cut -c1-400
echo "== ncu"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 100 -c 1 -o gpurun_out/prof_lean_c32 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lean_c32.log 2>&1; echo "rc=$?"
