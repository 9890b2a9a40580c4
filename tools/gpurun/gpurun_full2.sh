mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/e2e_profile.py 2>&1 | head -16 | tail -12
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 2500 gpurun_out/bench_r01.json
