#!/bin/bash
# First GPU call: smoke, parity tests, small + full bench, tuning sweep, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (without full-size)"; PR_SKIP_FULL=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 600 python bench.py --n-docs 2000000 --n-queries 8192 --steps 2 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "rc=$?"; cat gpurun_out/bench_small.json; tail -3 gpurun_out/bench_small.err
echo "== bench full"; timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== sweep"; timeout 1500 python tools/sweep.py --reps 1 > gpurun_out/sweep.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/sweep.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bm25 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_score -s 6 -c 2 -o gpurun_out/prof_score python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
