JJ_BENCH_INNER=3 JJ_BENCH_SKIP_E2E=1 timeout 600 compute-sanitizer --tool memcheck python bench.py --steps 1 --warmup 0 > gpurun_out/san.log 2>&1
grep -v "^$" gpurun_out/san.log | head -40
