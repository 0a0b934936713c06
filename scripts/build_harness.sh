#!/bin/bash
# Builds the standalone tensor-core parity/timing harness against the in-tree libmcn.so.
set -e
cd "$(dirname "$0")/.."
python -m myconvnet_b200.build
mkdir -p build
/usr/local/cuda/bin/nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
  tests/harness/tc_harness.cu -o build/tc_harness -Lmyconvnet_b200 -lmcn \
  -Xlinker -rpath -Xlinker '$ORIGIN/../myconvnet_b200'
echo built build/tc_harness
