mkdir -p gpurun_out
JJ_SUB_PROF=1 timeout 900 python tools/config_sweep.py cfg5 > gpurun_out/r2_prof3_cfg5.jsonl 2> gpurun_out/r2_prof3_cfg5.err
cut -c1-330 gpurun_out/r2_prof3_cfg5.jsonl
grep -v "sweep level\|stamp\|local cycles" gpurun_out/r2_prof3_cfg5.err | tail -80 | cut -c1-150
