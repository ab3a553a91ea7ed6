#!/bin/bash
# old (round-1 tree) vs new on a 1/8 shard and on the full corpus, one GPU
mkdir -p gpurun_out
for nd in 2626916 21015324; do
  echo "== r01 n_docs=$nd"; (cd build_variants/r01tree && timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c1-200)
  echo "== new n_docs=$nd"; timeout 600 python tools/latency.py --n-docs $nd --batches 65536 --reps 5 2>/dev/null | cut -c1-200
done
echo "== ncu r01 short"; (cd build_variants/r01tree && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_" --csv --log-file ../../gpurun_out/r2_ab_short_r01.csv python tools/latency.py --n-docs 2626916 --batches 65536 --reps 1 >/dev/null 2>&1)
echo "== ncu new short"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_" --csv --log-file gpurun_out/r2_ab_short_new.csv python tools/latency.py --n-docs 2626916 --batches 65536 --reps 1 >/dev/null 2>&1
wc -l gpurun_out/r2_ab_short_*.csv
