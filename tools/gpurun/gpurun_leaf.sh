for leaf in 8 12 24; do
JJ_LEAF_SIZE=$leaf JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_leaf.json 2> gpurun_out/subprof_leaf.err
echo "leaf $leaf"; grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_leaf.err | tail -21 | grep -v "sweep level" | grep -E "sweep|barrier 1|total"
done
