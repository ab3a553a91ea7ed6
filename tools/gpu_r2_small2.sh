#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/r2_pytest_bm25.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2_pytest_bm25.log
echo "== latency round0"; timeout 600 python tools/latency.py --batches 1,8,64,512,4096 --k 10 --reps 20 2>/dev/null | cut -c1-200
echo "== latency later"; timeout 600 python tools/latency.py --batches 1,8,64 --k 10 --reps 5 --kind later 2>/dev/null | cut -c1-200
