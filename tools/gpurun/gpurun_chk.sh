set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json
JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_auto.json 2> gpurun_out/subprof_auto.err
grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_auto.err | tail -31
