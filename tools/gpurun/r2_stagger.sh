mkdir -p gpurun_out
for half in 0 1; do
for stg in 0 6000 12000 24000 48000; do
JJ_SUB_STAGGER=$stg JJ_SUB_HALF=$half JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_stg.json 2> gpurun_out/r2_stg.err
python -c "
import json
d=json.load(open('gpurun_out/r2_stg.json')); print('half=$half stagger=$stg cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_stg.err
done
done
JJ_SUB_STAGGER=24000 JJ_SUB_HALF=0 JJ_SUB_PROF=1 JJ_BENCH_INNER=500 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_stg_prof.err
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_stg_prof.err | grep -v "sweep level" | tail -10 | cut -c1-160
