#!/bin/bash
# round-2 diagnostic: where the small-batch time goes (score vs merge kernels), before any change
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_diag1_smi.txt 2>&1
timeout 900 python tools/latency.py --batches 1,8,64,512,4096 --k 10 --reps 20 > gpurun_out/r2_lat_before.jsonl 2> gpurun_out/r2_lat_before.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|topk" --csv --log-file gpurun_out/r2_lat_launches.csv \
  python tools/latency.py --batches 1,8,64,512 --k 10 --reps 2 > gpurun_out/r2_lat_ncu.jsonl 2> gpurun_out/r2_lat_ncu.err
tail -5 gpurun_out/r2_lat_before.jsonl
