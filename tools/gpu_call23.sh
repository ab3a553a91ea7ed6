#!/bin/bash
# Vectorised fast producer + rem32 hot stream as defaults: smoke, parity suite, sweep, ncu.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c23.jsonl --configs "mode=6;mode=6,lazy_zero=2;mode=6,lazy_zero=1,subs_per_item=12;mode=6,subs_per_item=48,docs_per_launch=196608;mode=6,subs_per_item=24,docs_per_launch=98304,warps_per_cta=12" 2>&1 | grep -v "^\[bench" | cut -c1-300
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_flat -s 6 -c 1 -o gpurun_out/prof_flat_c23 python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_flat.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_flat.log | cut -c1-300
