#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu_c37.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_c37.log
echo "== sweep chunk-major"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c37.jsonl --configs "mode=8;mode=8,subs_per_item=12;mode=8,subs_per_item=8;mode=8,docs_per_launch=196608;mode=8,docs_per_launch=196608,subs_per_item=12;mode=8,docs_per_launch=393216,subs_per_item=12;mode=8,docs_per_launch=393216,subs_per_item=24;mode=8,docs_per_launch=786432,subs_per_item=16" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-400
