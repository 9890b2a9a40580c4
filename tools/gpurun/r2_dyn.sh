mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2_pytest_dyn.txt 2>&1
tail -6 gpurun_out/r2_pytest_dyn.txt
JJ_BENCH_SKIP_E2E=1 JJ_BENCH_CONFIGS=cfg3,cfg4,cfg5 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_dyn.json 2> gpurun_out/r2_dyn.err
python -c "
import json
d=json.load(open('gpurun_out/r2_dyn.json'))
print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))
for k,v in d['per_config'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','e2e','roofline_frac','device_us_per_time_step','setup_s','error')})
" || tail -5 gpurun_out/r2_dyn.err
JJ_SUB_PROF=1 timeout 900 python tools/config_sweep.py cfg4 > gpurun_out/r2_prof2_cfg4.jsonl 2> gpurun_out/r2_prof2_cfg4.err
grep -v "sweep level\|stamp\|local cycles\|upper phase" gpurun_out/r2_prof2_cfg4.err | tail -10 | cut -c1-150
