#!/bin/bash
mkdir -p gpurun_out
echo "== prober tests"; timeout 600 python -m pytest tests/test_gpu_prober.py -m gpu -x -q > gpurun_out/pytest_prober.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_prober.log
