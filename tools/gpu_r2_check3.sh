#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests/test_gpu_bm25.py -m gpu -x -q > gpurun_out/r2_pytest_bm25.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2_pytest_bm25.log
for nd in 2626916 21015324; do
  echo "== r01 n_docs=$nd"; (cd build_variants/r01tree && timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c60-140)
  echo "== new n_docs=$nd"; timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c60-140
done
