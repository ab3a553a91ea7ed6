#!/bin/bash
mkdir -p gpurun_out
for v in default $@ default $@; do
  if [ $v = default ]; then unset PR_LIB_PATH; else export PR_LIB_PATH=$PWD/build_variants/lib_$v.so; fi
  for nd in 2626916 21015324; do
    echo "== $v n_docs=$nd"; timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c60-140
  done
done
