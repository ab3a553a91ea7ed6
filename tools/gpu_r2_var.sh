#!/bin/bash
# variants (build_variants/lib_<name>.so) on the 1/8 shard and the full corpus, next to the round-1 tree
mkdir -p gpurun_out
echo "== r01"; (cd build_variants/r01tree && timeout 600 python tools/latency.py --n-docs 2626916 --batches 65536 --reps 5 2>/dev/null | cut -c60-140)
for v in default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  for nd in 2626916 21015324; do
    echo "== $v n_docs=$nd"; timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c60-140
  done
done
