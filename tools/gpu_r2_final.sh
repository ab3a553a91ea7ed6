#!/bin/bash
# Round-2 closing run (one GPU): whole GPU suite + smoke on the final build, 1M-passage text ingestion with the C
# front-end, prober at several batch sizes, the config-5 sweep and the default bench lines (both arms).
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
echo "== text build 1M"; timeout 900 python tools/text_build_bench.py --n 1000000 --out gpurun_out/r2f_text_build_1M.json 2>&1 | tail -n 2 | cut -c1-600
echo "== prober batch sizes"; rm -f gpurun_out/r2f_prober_batch_sizes.jsonl
for R in 256 1024 4096 16384 65536; do
  timeout 300 python tools/bench_prober.py --rows $R --out gpurun_out/_p.json > /dev/null 2>&1 && python -c "
import json; print(json.dumps(json.load(open('gpurun_out/_p.json'))))" >> gpurun_out/r2f_prober_batch_sizes.jsonl
done
rm -f gpurun_out/_p.json; wc -l gpurun_out/r2f_prober_batch_sizes.jsonl
bash tools/gpu_r2_evidence.sh sweep
bash tools/gpu_r2_evidence.sh bench
