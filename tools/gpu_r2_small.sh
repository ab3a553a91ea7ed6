#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/small_sweep.jsonl
timeout 900 python tools/small_batch_sweep.py --batches 1,8,64,512 --ipw 1,2,4,8,16,32 --nw 8,12 2>/dev/null | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_" --csv --log-file gpurun_out/r2_small_launches.csv python tools/latency.py --batches 1,8,64 --k 10 --reps 1 > /dev/null 2>&1
grep -c bm25 gpurun_out/r2_small_launches.csv
