mkdir -p gpurun_out
for dbg in 256 384; do
JJ_SUB_DEBUG=$dbg JJ_SUB_PROF=1 JJ_BENCH_INNER=100 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/itp.json 2> gpurun_out/itp.err
echo "dbg $dbg"; grep -A 40 "JJ_SUB_PROF" gpurun_out/itp.err | tail -22 | grep "junction\|face pass\|total"
done
tail -2 gpurun_out/itp.err
