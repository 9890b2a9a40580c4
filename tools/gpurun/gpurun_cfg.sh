mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
SWEEP_CHECK=1 timeout 1400 python tools/config_sweep.py cfg1 cfg3 cfg4 2> gpurun_out/cfg_sweep.err | tee gpurun_out/cfg_sweep.jsonl
