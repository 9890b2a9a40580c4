for v in 0 1 2 3; do
JJ_SUB_DEBUG=$v JJ_SUB_PROF=1 JJ_BENCH_INNER=100 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_auto.json 2> gpurun_out/subprof_auto.err
echo "debug $v"; grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_auto.err | tail -21 | grep "top product"
done
