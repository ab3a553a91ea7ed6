#!/bin/bash
for g in 1 2 3 6; do
  echo "== group $g"; PR_PROBER_GROUP=$g timeout 300 python tools/bench_prober.py --rows 16384 --out /dev/null 2>&1 | tail -n 1 | cut -c1-60
done
for rows in 1024 4096 65536; do echo "== rows $rows"; timeout 300 python tools/bench_prober.py --rows $rows --out /dev/null 2>&1 | tail -n 1 | cut -c1-160; done
