#!/bin/bash
mkdir -p gpurun_out
for v in default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python -m pytest tests/test_gpu_prober.py -m gpu -x -q 2>&1 | tail -n 1
  timeout 300 python tools/bench_prober.py --rows 16384 --out gpurun_out/r2_prober_bench_$v.json 2>&1 | tail -n 1 | cut -c1-120
done
