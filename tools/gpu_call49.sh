#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu_c49.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu_c49.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 3 --out gpurun_out/sweep_c49.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== sweep 1/8 shard"; timeout 1500 python tools/sweep.py --n-docs 2626916 --reps 5 --out gpurun_out/sweep_c49s.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
