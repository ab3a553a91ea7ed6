#!/bin/bash
# A/B of the posting-load ring depth / CTAs per SM for the DRAM-bound small-batch regime (variants built by tools/build_variant.sh)
mkdir -p gpurun_out
for v in default p3c3 p3c2 p4c2 p6c2; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  echo "== $v"
  timeout 600 python tools/latency.py --batches 1,8,64,512,4096,65536 --k 10 --reps 10 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/r2_pipe_ab.jsonl | cut -c1-200
done
