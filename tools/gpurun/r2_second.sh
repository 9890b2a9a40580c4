mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/r2_pytest.txt
cat gpurun_out/r2_pytest.txt
JJ_SUB_PROF=1 timeout 600 python tools/config_sweep.py cfg4 > gpurun_out/r2_prof4.jsonl 2> gpurun_out/r2_prof4.err
grep -A 14 "JJ_SUB_PROF" gpurun_out/r2_prof4.err | grep -v "sweep level" | tail -14
JJ_SUB_PROF=1 timeout 600 python tools/config_sweep.py cfg3 > gpurun_out/r2_prof3.jsonl 2> gpurun_out/r2_prof3.err
grep -A 14 "JJ_SUB_PROF" gpurun_out/r2_prof3.err | grep -v "sweep level" | tail -14
