#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu bm25"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu_c35.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_c35.log
echo "== sweep main (sign epochs + free-bank pads)"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c35.jsonl --configs "mode=8;mode=8,warps_per_cta=10;mode=6" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-400
echo "== sweep ns (free-bank pads only)"; PR_LIB_PATH=$PWD/build_variants/lib_ns.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c35_ns.jsonl --configs "mode=8;mode=8,warps_per_cta=10" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300
