#!/bin/bash
# prober bench (BASELINE config 4 shape) + ncu --set full of its kernels
mkdir -p gpurun_out
timeout 600 python tools/bench_prober.py --rows 16384 --out gpurun_out/prober_bench.json 2>&1 | tail -1 | cut -c1-800
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prober_ -s 6 -c 6 -o gpurun_out/prof_prober python tools/bench_prober.py --rows 16384 --out gpurun_out/prober_bench_ncu.json > gpurun_out/ncu_prober.log 2>&1; echo "rc=$?"
