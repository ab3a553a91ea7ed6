#!/bin/bash
# Round-1 evidence run: full GPU suite, default bench, ncu launch list of the bench command, ncu --set full of the scoring kernel.
mkdir -p gpurun_out
echo "== pytest gpu (all, incl. full-size)"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu_full.log
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/bench_c26.json 2> gpurun_out/bench_c26.err; echo "rc=$?"; cut -c1-2500 gpurun_out/bench_c26.json
echo "== ncu launch list"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c26.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_c26.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_c26.csv
echo "== ncu full (21M docs, launch in the middle of a step)"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_flat -s 100 -c 1 -o gpurun_out/prof_flat_c26 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_flat_c26.log 2>&1; echo "rc=$?"
