#!/bin/bash
# N-GPU checks: NCCL parity test, bench with / without the threshold exchange.  usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
echo "== pytest multirank"; timeout 1500 python -m pytest tests/test_gpu_multirank.py tests/test_training.py -m gpu -x -q > gpurun_out/r2_pytest_multi_$N.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2_pytest_multi_$N.log
for mode in p2p allreduce none; do
  extra="--exchange $mode"
  echo "== bench N=$N $mode"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 $extra > gpurun_out/r2_bench_n${N}_$mode.json 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "rc=$?"
  python - <<PY
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','result_digest')}, 'e2e', round(l['e2e']['value']), 'kernel ms', l['roofline']['kernel_ms_per_step'], 'frac', l['roofline']['frac'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/r2_bench_n${N}_$mode.err').read()[-1500:])
PY
done
