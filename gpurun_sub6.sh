JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_auto.json 2> gpurun_out/subprof_auto.err
grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_auto.err | tail -21 | grep -v "sweep level"
timeout 600 python -m pytest tests -m gpu -x -q -k "subdomain_solve" 2>&1 | tail -2
