mkdir -p gpurun_out
JJ_SUB_HALF=1 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "(subdomain_engine_matches or subdomain_solve_matches or full_size_cfg2 or device_matches_reference_golden or running_observables_equal) and not upper" > gpurun_out/r2_half_tests.txt 2>&1
tail -5 gpurun_out/r2_half_tests.txt
for half in 0 1; do
JJ_SUB_HALF=$half JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_half_$half.json 2> gpurun_out/r2_half_$half.err
python -c "
import json
d=json.load(open('gpurun_out/r2_half_$half.json')); print('half=$half cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']), d['config']['subdomains_or_cluster'], d['config']['problems_per_block'])" || tail -5 gpurun_out/r2_half_$half.err
done
JJ_SUB_HALF=1 JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_half_prof.err
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_half_prof.err | grep -v "sweep level" | tail -12 | cut -c1-160
