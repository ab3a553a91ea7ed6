#!/bin/bash
# one query per call: mean device time over 64 distinct queries, and the per-kernel split (ncu launch list)
mkdir -p gpurun_out
timeout 600 python tools/latency.py --batches 1,8,64 --k 1,10,100 --reps 10 --graph 2>/dev/null | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_" --csv --log-file gpurun_out/r2_single_launches.csv python tools/latency.py --batches 1 --k 10 --reps 1 > /dev/null 2>&1
grep -c bm25 gpurun_out/r2_single_launches.csv
