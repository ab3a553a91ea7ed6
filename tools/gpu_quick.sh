#!/bin/bash
# quick check of a kernel change: BM25 parity suite + full-size sweep of the default mode (+ extra configs in $1)
mkdir -p gpurun_out
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu_quick.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu_quick.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 3 --out gpurun_out/sweep_quick.jsonl --configs "mode=8${1:+;$1}" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
