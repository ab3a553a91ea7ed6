#!/bin/bash
# same-box A/B: the in-tree library against build_variants/lib_<name>.so (tools/build_variant.sh): bm25 tests, full corpus + 1/8 shard + small batches; usage: gpu_r2_ab.sh name...
mkdir -p gpurun_out
echo "== bm25 tests (in-tree build)"; timeout 900 python -m pytest tests/test_gpu_bm25.py tests/test_gpu_retriever.py -x -q 2>&1 | tail -n 3
for v in default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  for nd in 2626916 21015324; do
    echo "== $v n_docs=$nd"; timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c60-140
  done
  echo "== $v small"; timeout 600 python tools/latency.py --batches 8,64,512,4096 --k 10 --reps 10 2>/dev/null | cut -c1-140
done
