mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "anneal" > gpurun_out/r2_pytest_anneal.txt 2>&1
tail -12 gpurun_out/r2_pytest_anneal.txt | cut -c1-200
timeout 600 python tools/anneal_bench.py > gpurun_out/r2_anneal_bench.json 2> gpurun_out/r2_anneal_bench.err
tail -3 gpurun_out/r2_anneal_bench.err; cat gpurun_out/r2_anneal_bench.json | cut -c1-1500
JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_it3.json 2> gpurun_out/r2_it3.err
python -c "
import json
d=json.load(open('gpurun_out/r2_it3.json')); print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f e2e %.2f G (%.1f ms)'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['e2e']['seconds_per_step']*1e3))" || tail -5 gpurun_out/r2_it3.err
