mkdir -p gpurun_out
for pol in 00 11 22 20 02 12 21; do
JJ_SUB_STATE_POL=$pol JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/pol_$pol.json 2> gpurun_out/pol_$pol.err
python -c "
import json
d=json.load(open('gpurun_out/pol_$pol.json')); print('pol $pol us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/pol_$pol.err
done
