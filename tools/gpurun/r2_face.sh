mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "golden or full_size or running_observables or larger_circuit or vortex" > gpurun_out/r2_pytest_face.txt 2>&1
tail -4 gpurun_out/r2_pytest_face.txt
JJ_BENCH_SKIP_E2E=1 JJ_BENCH_CONFIGS=cfg3,cfg4 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_face.json 2> gpurun_out/r2_face.err
python -c "
import json
d=json.load(open('gpurun_out/r2_face.json'))
print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))
for k,v in d['per_config'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','roofline_frac','device_us_per_time_step','error')})
" || tail -5 gpurun_out/r2_face.err
JJ_SUB_PROF=1 JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_face_prof.err
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_face_prof.err | grep -v "sweep level" | tail -10 | cut -c1-160
