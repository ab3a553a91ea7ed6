#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (without full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_warp2.jsonl --configs "mode=4,warps_per_cta=8;mode=4,warps_per_cta=13;mode=4,warps_per_cta=4;mode=4,warps_per_cta=8,lazy_zero=2;mode=3,warps_per_cta=8;mode=4,warps_per_cta=8,subs_per_item=24,docs_per_launch=196608" > gpurun_out/sweep_warp2.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/sweep_warp2.log
bash tools/gpu_prof.sh mode=4,warps_per_cta=8 prof_warp5
