#!/bin/bash
# prober parity + timing (own short timeouts: a wrong barrier protocol hangs the kernel)
mkdir -p gpurun_out
echo "== pytest prober"; timeout 600 python -m pytest tests/test_gpu_prober.py -m gpu -x -q > gpurun_out/r2_pytest_prober.log 2>&1; echo "rc=$?"; tail -n 8 gpurun_out/r2_pytest_prober.log
echo "== bench prober"; timeout 300 python tools/bench_prober.py --rows 16384 --out gpurun_out/r2_prober_bench.json 2>&1 | tail -n 3; cat gpurun_out/r2_prober_bench.json 2>/dev/null | cut -c1-600
echo; echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"prober_" -c 20 --csv --log-file gpurun_out/r2_prober_launches.csv python tools/bench_prober.py --rows 16384 --out /dev/null > /dev/null 2>&1; grep -c prober gpurun_out/r2_prober_launches.csv
