#!/bin/bash
# Build libjjstep.so in-tree for sm_100a (cross-compiles without a GPU). Translation units are compiled in parallel;
# the step kernel of the subdomain engine is one unit per chunk width (jj_subdomain.cu with -DJJ_SUB_NG=...).
# JJ_NVCC_EXTRA adds flags (e.g. -DJJ_EXPERIMENTS for the timing experiments of the profiling notes).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# --register-usage-level=2: ptxas holds back the optimisations that trade registers for speed; the step kernel lives at
# its 128-register cap, and this takes the lean kernel from 112 to 12 bytes of spills (cfg2 +0.2 %, cfg3 +1.1 %, cfg4
# +1.6 %, measured A/B on one B200, profiles/r02_experiments.md)
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -Xptxas --register-usage-level=2 $JJ_NVCC_EXTRA"
mkdir -p obj
pids=()
names=()
run() {  # name, args...
    local name=$1; shift
    ( $NVCC $FLAGS -c "$@" -o obj/$name.o > obj/$name.log 2>&1 ) &
    pids+=($!); names+=($name)
}
run jjstep jjstep.cu
run jj_observe jj_observe.cu
run jj_subdomain jj_subdomain.cu
for ng in 1 2 4 8; do run jj_subdomain_ng$ng -DJJ_SUB_NG=$ng jj_subdomain.cu; done
for ng in 1 2 4; do run jj_subdomain_h_ng$ng -DJJ_SUB_NG=$ng -DJJ_SUB_NT=256 jj_subdomain.cu; done
for ng in 1 2 4 8; do run jj_subdomain_m_ng$ng -DJJ_SUB_NG=$ng -DJJ_SUB_MOB=1 jj_subdomain.cu; done
fail=0
for i in "${!pids[@]}"; do
    if ! wait ${pids[$i]}; then echo "== ${names[$i]} failed"; cat obj/${names[$i]}.log; fail=1; fi
done
cat obj/*.log > build.log
[ $fail = 0 ] || exit 1
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ${JJ_LIB_OUT:-../libjjstep.so} obj/*.o
grep -E "error|warning" build.log | grep -v "^ptxas info" || true
echo "built $(cd ..; pwd)/libjjstep.so"
