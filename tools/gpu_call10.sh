#!/bin/bash
mkdir -p gpurun_out
echo "== prober bench"; timeout 600 python tools/bench_prober.py 2>&1 | tail -2
echo "== prober ncu"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:prober_ -s 30 -c 4 -o gpurun_out/prof_prober python tools/bench_prober.py --rows 16384 > gpurun_out/prof_prober.log 2>&1; echo "rc=$?"
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== full gpu tests incl. 21M"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_all.log
echo "== bench default"; timeout 1500 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?"; cat gpurun_out/bench_r01.json | cut -c1-1500; tail -2 gpurun_out/bench_r01.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-900
