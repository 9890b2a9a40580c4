mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "anneal or mobility or zone or golden" > gpurun_out/r2_pytest_anneal.txt 2>&1
tail -3 gpurun_out/r2_pytest_anneal.txt
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
ANNEAL_LOOP=0 timeout 600 python tools/anneal_bench.py 2>/dev/null | head -c 420; echo
JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2 us/timestep %.2f frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['roofline']['frac']))"
