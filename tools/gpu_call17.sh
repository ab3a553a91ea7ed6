#!/bin/bash
# Mode 7 (rank-safe term skipping + exact rescoring): parity suite, then mode sweep at 21M.
mkdir -p gpurun_out
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c17.jsonl --configs "mode=6,warps_per_cta=8;mode=7,warps_per_cta=8;mode=7,warps_per_cta=8,rescore_cost=16;mode=7,warps_per_cta=8,rescore_cost=256;mode=7,warps_per_cta=8,subs_per_item=24;mode=7,warps_per_cta=12;mode=7,warps_per_cta=4;mode=7,warps_per_cta=8,docs_per_launch=49152,subs_per_item=12" 2>&1 | grep -v "^\[bench" | cut -c1-400
echo "== ncu full skip"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_flat -s 6 -c 1 -o gpurun_out/prof_skip_c17 python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline --tune mode=7 > gpurun_out/ncu_skip.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_skip.log | cut -c1-300
