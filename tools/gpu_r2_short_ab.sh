#!/bin/bash
# per-launch times of one 64k-query call over a 1/8 shard: in-tree library against build_variants/lib_<name>.so
mkdir -p gpurun_out
for v in default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_lean" --csv --log-file gpurun_out/r2_short_launches_$v.csv python tools/latency.py --n-docs 2626916 --batches 65536 --reps 1 > gpurun_out/r2_short_$v.log 2>&1
  tail -n 1 gpurun_out/r2_short_$v.log | cut -c1-200
done
