mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for dbg in 0; do
JJ_SUB_DEBUG=$dbg JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/it.json 2> gpurun_out/it.err
python -c "
import json
d=json.load(open('gpurun_out/it.json')); print('dbg $dbg us/timestep %.1f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/it.err
done
JJ_SUB_PROF=1 JJ_BENCH_INNER=200 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 1 --warmup 1 > gpurun_out/itp.json 2> gpurun_out/itp.err
grep -A 30 "JJ_SUB_PROF" gpurun_out/itp.err | tail -18 | grep -v "sweep level"
