#!/bin/bash
# Evidence run for the lean kernel default: bench line, ncu launch list of the bench command, ncu --set full capture.
mkdir -p gpurun_out
echo "== bench default"; timeout 1200 python bench.py > gpurun_out/bench_c45.json 2> gpurun_out/bench_c45.err; echo "rc=$?"; cut -c1-3000 gpurun_out/bench_c45.json
echo "== ncu launch list (our kernels only)"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bm25_ -c 1300 --csv --log-file gpurun_out/launches_c45.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_c45.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_c45.csv
echo "== ncu full"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:bm25_lean -s 60 -c 1 -o gpurun_out/prof_lean_c45 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lean_c45.log 2>&1; echo "rc=$?"
