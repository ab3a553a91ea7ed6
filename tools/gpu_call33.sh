#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu bm25 + prober"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_prober.py -m gpu -x -q > gpurun_out/pytest_gpu_c33.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_c33.log
echo "== sweep main"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c33.jsonl --configs "mode=8;mode=8,warps_per_cta=10;mode=8,docs_per_launch=49152;mode=8,docs_per_launch=49152,subs_per_item=12;mode=8,warps_per_cta=10,docs_per_launch=49152;mode=8,docs_per_launch=196608,subs_per_item=24" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-400
for v in p2 p4; do echo "== sweep $v"; PR_LIB_PATH=$PWD/build_variants/lib_$v.so timeout 900 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c33_$v.jsonl --configs "mode=8;mode=8,warps_per_cta=10" 2>&1 | grep -v "^\[bench" | grep -v aux | cut -c1-300; done
