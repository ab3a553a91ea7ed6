#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== pytest gpu (all)"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c51.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_c51.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 3 --out gpurun_out/sweep_c51.jsonl --configs "mode=8;mode=6;mode=4" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
echo "== sweep 1/8 shard"; timeout 1500 python tools/sweep.py --n-docs 2626916 --reps 5 --out gpurun_out/sweep_c51s.jsonl --configs "mode=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
