#!/bin/bash
# Build libjjstep.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
$NVCC $FLAGS $JJ_NVCC_EXTRA -shared -o ${JJ_LIB_OUT:-../libjjstep.so} jjstep.cu jj_resident.cu jj_subdomain.cu jj_observe.cu 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "^ptxas info" || true
echo "built $(cd ..; pwd)/libjjstep.so"
