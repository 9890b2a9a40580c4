for cfg in 4,4; do
  JJ_SUB_PROF=1 JJ_SUBDOMAIN=$cfg JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_$cfg.json 2> gpurun_out/subprof_$cfg.err
  grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_$cfg.err | tail -32
done
