mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2_pytest_obs.txt 2>&1
tail -15 gpurun_out/r2_pytest_obs.txt
JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_it2.json 2> gpurun_out/r2_it2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_it2.json')); print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f e2e %.2f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9))" || tail -5 gpurun_out/r2_it2.err
for c in cfg4 cfg3; do
timeout 600 python tools/e2e_profile_cfg.py $c 200 > gpurun_out/r2_e2eprof2_$c.txt 2>&1
head -22 gpurun_out/r2_e2eprof2_$c.txt | cut -c1-150
done
