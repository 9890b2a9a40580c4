mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "subdomain_solve or golden or full_size_cfg2" > gpurun_out/r2_pytest_quick.txt 2>&1
tail -3 gpurun_out/r2_pytest_quick.txt
JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_quick.json 2> gpurun_out/r2_quick.err
python -c "
import json
d=json.load(open('gpurun_out/r2_quick.json')); print('cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_quick.err
JJ_SUB_PROF=1 JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_quick_prof.err
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_quick_prof.err | grep -v "sweep level" | tail -9 | cut -c1-160
