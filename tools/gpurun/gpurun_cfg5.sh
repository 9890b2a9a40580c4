mkdir -p gpurun_out
timeout 1700 python tools/config_sweep.py cfg5 2> gpurun_out/cfg_sweep5.err | tee gpurun_out/cfg_sweep5.jsonl
tail -3 gpurun_out/cfg_sweep5.err
