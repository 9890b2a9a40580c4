JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_auto.json 2> gpurun_out/subprof_auto.err
grep -A 30 "JJ_SUB_PROF" gpurun_out/subprof_auto.err | tail -21 | grep -v "sweep level"
JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/sub_auto.json 2> gpurun_out/sub_auto.err
python -c "
import json
d=json.load(open('gpurun_out/sub_auto.json')); print('auto us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/sub_auto.err
