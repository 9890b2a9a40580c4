mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/r2_pytest.txt
cat gpurun_out/r2_pytest.txt
for c in cfg4 cfg5; do
JJ_SUB_PROF=1 timeout 900 python tools/config_sweep.py $c > gpurun_out/r2_prof_$c.jsonl 2> gpurun_out/r2_prof_$c.err
cat gpurun_out/r2_prof_$c.jsonl
grep -A 200 "JJ_SUB_PROF" gpurun_out/r2_prof_$c.err | grep -v "sweep level" | tail -120 | grep "upper\|JJ_SUB" 
done
