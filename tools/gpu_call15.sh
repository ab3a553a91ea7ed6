#!/bin/bash
# Re-validation after container re-creation: smoke, GPU parity suite, mode sweep at 21M, prober bench.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu (without full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 2 --out gpurun_out/sweep_c15.jsonl --configs "mode=4,warps_per_cta=8;mode=6,warps_per_cta=8;mode=6,warps_per_cta=12;mode=6,warps_per_cta=4;mode=5,warps_per_cta=8;mode=6,warps_per_cta=8,subs_per_item=24;mode=6,warps_per_cta=8,docs_per_launch=65536,subs_per_item=8;mode=6,warps_per_cta=8,docs_per_launch=196608,subs_per_item=24" 2>&1 | grep -v "^\[bench" | cut -c1-400
echo "== prober bench"; timeout 600 python tools/bench_prober.py 2>&1 | tail -5
echo "== ncu full flat"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_flat -s 6 -c 1 -o gpurun_out/prof_flat_c15 python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline --tune mode=6 > gpurun_out/ncu_flat.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_flat.log | cut -c1-300
