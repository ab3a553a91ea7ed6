#!/bin/bash
# union-bound exchange A/B at N GPUs: bench.py with --list-rounds R for every R given; usage: gpu_r2_union.sh N R...
N=${1:-2}; shift; ROUNDS=${@:-0 4}
mkdir -p gpurun_out
if [ $N = 2 ]; then echo "== pytest multirank"; timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -n 3; fi
for r in $ROUNDS; do
  echo "== bench N=$N list-rounds=$r"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --list-rounds $r > gpurun_out/r2u_bench_n${N}_lr$r.json 2> gpurun_out/r2u_bench_n${N}_lr$r.err; echo "rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2u_bench_n${N}_lr$r.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','result_digest')}, 'e2e', round(l['e2e']['value']), 'kernel ms', l['roofline']['kernel_ms_per_step'], 'frac', l['roofline']['frac'], l['plan'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/r2u_bench_n${N}_lr$r.err').read()[-1500:])
PY
done
