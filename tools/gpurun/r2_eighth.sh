mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/r2_pytest_full.txt 2>&1
tail -3 gpurun_out/r2_pytest_full.txt; grep -n "Fatal\|Segmentation\|Timeout\|Error" gpurun_out/r2_pytest_full.txt | head
JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_it.json 2> gpurun_out/r2_it.err
python -c "
import json
d=json.load(open('gpurun_out/r2_it.json')); print('cfg2 us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_it.err
for c in cfg4; do
JJ_SUB_PROF=1 timeout 900 python tools/config_sweep.py $c > gpurun_out/r2_prof_$c.jsonl 2> gpurun_out/r2_prof_$c.err
cut -c1-250 gpurun_out/r2_prof_$c.jsonl
grep "stamp" gpurun_out/r2_prof_$c.err | tail -6
grep "upper phase  0" gpurun_out/r2_prof_$c.err | tail -2
done
