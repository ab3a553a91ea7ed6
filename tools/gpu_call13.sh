#!/bin/bash
# Flat-step kernel with the hot posting stream: parity tests, sweeps, ncu.
mkdir -p gpurun_out
echo "== pytest gpu bm25 (without full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== sweep hot"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_hot.jsonl --configs "mode=4,warps_per_cta=8;mode=6,warps_per_cta=8;mode=6,warps_per_cta=12;mode=6,warps_per_cta=4;mode=5,warps_per_cta=8;mode=6,warps_per_cta=8,subs_per_item=24;mode=6,warps_per_cta=8,docs_per_launch=65536,subs_per_item=8;mode=6,warps_per_cta=8,docs_per_launch=49152,subs_per_item=12" 2>&1 | grep -v "^\[bench" | cut -c1-300
for v in pipe2 pipe4 occ2; do
echo "== $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_$v.jsonl --configs "mode=6,warps_per_cta=8" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-260
done
echo "== ncu full flat"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_flat -s 6 -c 1 -o gpurun_out/prof_flat_v2 python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline --tune mode=6 > gpurun_out/ncu_flat.log 2>&1; echo "rc=$?"
