mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "vortex or observ or mobility" > gpurun_out/r2_pytest_vortex.txt 2>&1
tail -3 gpurun_out/r2_pytest_vortex.txt
timeout 600 python tools/vortex_tracking_bench.py > gpurun_out/r2_vortex_tracking.json 2> gpurun_out/r2_vortex_tracking.err; tail -2 gpurun_out/r2_vortex_tracking.err; cat gpurun_out/r2_vortex_tracking.json
