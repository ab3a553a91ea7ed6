#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all, incl. full-size)"; timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c41.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu_c41.log
