#!/bin/bash
# parity suite of the BM25 path + small-batch latency + full-size step
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_retriever.py -m gpu -x -q > gpurun_out/r2_pytest_bm25.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/r2_pytest_bm25.log
echo "== latency"; timeout 600 python tools/latency.py --batches 1,8,64,512,4096,65536 --k 10 --reps 10 2>gpurun_out/r2_lat.err | tee gpurun_out/r2_lat_after.jsonl | cut -c1-220
echo "== ncu launch list 64k"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_" --csv --log-file gpurun_out/r2_launches_64k.csv python tools/latency.py --batches 65536 --k 10 --reps 1 > /dev/null 2>&1; wc -l gpurun_out/r2_launches_64k.csv
