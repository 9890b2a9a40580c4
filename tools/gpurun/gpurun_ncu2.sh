set -x
mkdir -p gpurun_out
JJ_BENCH_INNER=10 JJ_BENCH_SKIP_E2E=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subdomain -s 1 -c 1 -o gpurun_out/prof_sub_v3 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_sub_v3.log 2>&1
tail -5 gpurun_out/ncu_sub_v3.log
ls -la gpurun_out/
