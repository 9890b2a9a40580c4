timeout 900 python -m pytest tests -m gpu -x -q -k "subdomain" 2>&1 | tail -5
for cfg in 4,4 5,8; do
  JJ_SUB_PROF=1 JJ_SUBDOMAIN=$cfg JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/subprof_$cfg.json 2> gpurun_out/subprof_$cfg.err
  tail -10 gpurun_out/subprof_$cfg.err
  JJ_SUBDOMAIN=$cfg JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 2 --warmup 1 > gpurun_out/sub_$cfg.json 2> gpurun_out/sub_$cfg.err
  python -c "
import json
d=json.load(open('gpurun_out/sub_$cfg.json')); print('cfg $cfg us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/sub_$cfg.err
done
