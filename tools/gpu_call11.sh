#!/bin/bash
# Re-validation after container re-creation: smoke, GPU parity tests, default bench, prober bench, ncu.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (without full-size)"; PR_SKIP_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== prober bench"; timeout 600 python tools/bench_prober.py 2>&1 | tail -2
echo "== bench default"; timeout 1500 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_r01.json; tail -3 gpurun_out/bench_r01.err
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --n-docs 2000000 --n-queries 65536 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_warp -s 6 -c 2 -o gpurun_out/prof_warp_r01 python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
