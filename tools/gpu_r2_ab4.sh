#!/bin/bash
# in-tree library against build_variants/lib_<name>.so: bm25 tests on the in-tree build, then full corpus + 1/8 shard
mkdir -p gpurun_out
echo "== bm25 tests (in-tree build)"; timeout 600 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_retriever.py -x -q 2>&1 | tail -n 3
for v in default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  echo "== $v"; timeout 300 python tools/latency.py --batches 64,4096,65536 --reps 5 2>/dev/null | cut -c1-150
done
