mkdir -p gpurun_out
JJ_LIB_PATH=$PWD/pyjjasim_b200/libjjstep_w32.so JJ_SUB_WARPS=32 JJ_SUB_PROF=1 JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/r2_w32.json 2> gpurun_out/r2_w32.err
tail -3 gpurun_out/r2_w32.err | cut -c1-300
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_w32.err | tail -22 | cut -c1-160
