set -x
mkdir -p gpurun_out
JJ_BENCH_INNER=10 JJ_BENCH_SKIP_E2E=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subdomain -s 1 -c 1 -o gpurun_out/prof_sub_v4 python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_sub_v4.log 2>&1
tail -3 gpurun_out/ncu_sub_v4.log
JJ_BENCH_INNER=50 JJ_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_r01.log 2>&1
tail -12 gpurun_out/launches_r01.csv
