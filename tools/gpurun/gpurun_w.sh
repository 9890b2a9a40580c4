mkdir -p gpurun_out
for w in 256 128 32; do
JJ_BENCH_W=$w JJ_BENCH_NPARTS=18 JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/w_$w.json 2> gpurun_out/w_$w.err
echo "=== W=$w"; grep -A 30 "JJ_SUB_PROF" gpurun_out/w_$w.err | tail -18 | grep -v "sweep level"
done
