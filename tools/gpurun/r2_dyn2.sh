mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2_pytest_dyn2.txt 2>&1
tail -4 gpurun_out/r2_pytest_dyn2.txt
JJ_BENCH_SKIP_E2E=1 JJ_BENCH_CONFIGS=cfg3,cfg4 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_dyn2.json 2> gpurun_out/r2_dyn2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_dyn2.json'))
print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))
for k,v in d['per_config'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','e2e','roofline_frac','device_us_per_time_step','setup_s','error')})
" || tail -5 gpurun_out/r2_dyn2.err
