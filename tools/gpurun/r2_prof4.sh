mkdir -p gpurun_out
JJ_SUB_PROF=1 timeout 900 python tools/config_sweep.py cfg4 > gpurun_out/r2_prof4_cfg4.jsonl 2> gpurun_out/r2_prof4_cfg4.err
cut -c1-330 gpurun_out/r2_prof4_cfg4.jsonl
grep "stamp [567]\|bwd sweep\|fwd sweep\|sweep level" gpurun_out/r2_prof4_cfg4.err | tail -14 | cut -c1-250
