#!/bin/bash
# whole gpu suite + smoke + default bench line
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"; tail -n 3 gpurun_out/r2_bench_n1.err; python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches','result_digest','sample_digest','parity_checked')})
    print('e2e',l['e2e']); print('roofline frac',l['roofline']['frac'],'kernel ms',l['roofline']['kernel_ms_per_step'])
    print('cpu',[(v['name'],round(v['value'],2),v['cores'],v['sample'][:12]) for v in l['cpu_baseline']['variants']])
    print('secondary c4',l['secondary'].get('config4_prober_gated_batch') or l['secondary'])
    for r in l['secondary'].get('config5_sweep',[]): print(r)
except Exception as e: print('parse failed',e)
PY
