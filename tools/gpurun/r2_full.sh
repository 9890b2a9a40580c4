mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2_pytest_full2.txt 2>&1
tail -6 gpurun_out/r2_pytest_full2.txt | cut -c1-200
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_full2.json 2> gpurun_out/r2_bench_full2.err
tail -4 gpurun_out/r2_bench_full2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_full2.json'))
print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f e2e %.2f G'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9))
for k,v in d['per_config'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','e2e','roofline_frac','device_us_per_time_step','setup_s','error')})
print(d['cpu_baseline']['value'])
"
timeout 600 python tools/anneal_bench.py > gpurun_out/r2_anneal_bench.json 2> gpurun_out/r2_anneal_bench.err; tail -2 gpurun_out/r2_anneal_bench.err; head -c 500 gpurun_out/r2_anneal_bench.json; echo
timeout 300 python __graft_entry__.py > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
JJ_ANNEAL_INTERVALS=6 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/anneal_launches.csv python tools/anneal_launches.py > gpurun_out/anneal_launches.log 2>&1
tail -1 gpurun_out/anneal_launches.log
