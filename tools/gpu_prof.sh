#!/bin/bash
# ncu full capture of the scoring kernel (args: tuning string, output name)
mkdir -p gpurun_out
TUNE=${1:-mode=4}
NAME=${2:-prof}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bm25_ -s 6 -c 2 -o gpurun_out/$NAME python bench.py --n-docs 2000000 --n-queries 65536 --steps 1 --warmup 1 --no-cpu-baseline --tune $TUNE > gpurun_out/${NAME}.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${NAME}.log | cut -c1-600
