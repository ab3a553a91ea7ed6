#!/bin/bash
# final scaling line at N GPUs (p2p threshold exchange); usage: gpu_r2_scale.sh N [modes...]
N=${1:-2}; shift; MODES=${@:-p2p}
mkdir -p gpurun_out
if [ $N = 2 ]; then echo "== pytest multirank"; timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -n 2; fi
for mode in $MODES; do
  echo "== bench N=$N $mode"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $mode > gpurun_out/r2f_bench_n${N}_$mode.json 2> gpurun_out/r2f_bench_n${N}_$mode.err; echo "rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2f_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','result_digest')}, 'e2e', round(l['e2e']['value']), 'kernel ms', l['roofline']['kernel_ms_per_step'], 'frac', l['roofline']['frac'], l['plan']['threshold_exchange'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/r2f_bench_n${N}_$mode.err').read()[-1500:])
PY
done
