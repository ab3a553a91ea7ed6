#!/bin/bash
# Hot-stream remainder variants (one padded wide step instead of several narrow ones), then the default bench.
mkdir -p gpurun_out
for v in rem32 rem64 rem32p4; do
echo "== $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_$v.jsonl --configs "mode=6,warps_per_cta=8;mode=6,warps_per_cta=8,subs_per_item=24" 2>&1 | grep -v "^\[bench" | cut -c1-300
done
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/bench_c20.json 2> gpurun_out/bench_c20.err; echo "rc=$?"; cut -c1-1500 gpurun_out/bench_c20.json; tail -3 gpurun_out/bench_c20.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c20.json 2> gpurun_out/bench_ref_c20.err; echo "rc=$?"; cut -c1-800 gpurun_out/bench_ref_c20.json
