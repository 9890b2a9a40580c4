mkdir -p gpurun_out
JJ_BENCH_INNER=10 JJ_BENCH_SKIP_E2E=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subdomain -s 1 -c 1 -o gpurun_out/prof_sub_final python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_sub_final.log 2>&1
tail -2 gpurun_out/ncu_sub_final.log
JJ_BENCH_INNER=50 JJ_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_final.log 2>&1
grep -c k_sub gpurun_out/launches_final.csv
