#!/bin/bash
# build_variants/lib_<name>.so with extra -D flags (A/B builds for tools/sweep.py via PR_LIB_PATH)
# usage: tools/build_variant.sh name "-DPR_LEAN_PIPE=4 ..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $2 \
  -o build_variants/lib_$1.so probing_rag_b200/csrc/*.cu -lcuda
echo built build_variants/lib_$1.so
