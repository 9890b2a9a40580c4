set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; tail -c 1500 gpurun_out/bench_r01_ref.json
JJ_BENCH_INNER=50 JJ_BENCH_SKIP_E2E=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_r01.log 2>&1
tail -5 gpurun_out/launches_r01.csv
