#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (without full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== sweep warp"; timeout 1500 python tools/sweep.py --reps 1 --grid warp --out gpurun_out/sweep_warp.jsonl > gpurun_out/sweep_warp.log 2>&1; echo "rc=$?"; tail -45 gpurun_out/sweep_warp.log
