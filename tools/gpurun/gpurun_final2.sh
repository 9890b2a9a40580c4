mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python tools/config_sweep.py cfg1 2>/dev/null | cut -c1-330
JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/it.json 2> gpurun_out/it.err
python -c "
import json
d=json.load(open('gpurun_out/it.json')); print('cfg2 us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/it.err
