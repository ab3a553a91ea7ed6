#!/bin/bash
# Lean-step kernel (mode 8): smoke, parity suite, full-size sweep against the flat kernel (mode 6).
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu_c30.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu_c30.log
echo "== sweep"; date; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c30.jsonl --configs "mode=8;mode=6;mode=8,warps_per_cta=10;mode=8,warps_per_cta=12;mode=8,subs_per_item=12;mode=8,docs_per_launch=65536,subs_per_item=16" 2>&1 | grep -v "^\[bench" | cut -c1-400; date
