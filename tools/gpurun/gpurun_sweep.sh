python -m pytest tests -m gpu -x -q -k "resident" 2>&1 | tail -5
for cfg in 4,8 8,8; do
  JJ_RESIDENT=$cfg JJ_BENCH_INNER=500 JJ_BENCH_SKIP_E2E=1 python bench.py --steps 2 --warmup 1 > gpurun_out/t.json 2> gpurun_out/t.err
  python -c "
import json
d=json.load(open('gpurun_out/t.json')); print('cfg $cfg us/timestep %.1f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step']))" || tail -3 gpurun_out/t.err
done
for cfg in 4,8 8,8; do
JJ_RES_PROF=1 JJ_RESIDENT=$cfg JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 python bench.py --steps 1 --warmup 1 > gpurun_out/prof.json 2> gpurun_out/prof.err
grep -A 60 "JJ_RES_PROF" gpurun_out/prof.err | tail -58 > gpurun_out/prof_$cfg.txt; head -3 gpurun_out/prof_$cfg.txt; tail -1 gpurun_out/prof_$cfg.txt
done
