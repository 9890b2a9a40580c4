# round 2, first GPU call: parity suite after the engine refactor (upper program), short bench, cfg sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 > gpurun_out/r2_pytest.txt
cat gpurun_out/r2_pytest.txt
JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_it.json 2> gpurun_out/r2_it.err
python -c "
import json
d=json.load(open('gpurun_out/r2_it.json')); print('cfg2 us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_it.err
timeout 900 python tools/config_sweep.py cfg4 cfg3 > gpurun_out/r2_sweep.jsonl 2> gpurun_out/r2_sweep.err
cat gpurun_out/r2_sweep.jsonl; tail -3 gpurun_out/r2_sweep.err
