mkdir -p gpurun_out
for leaf in 24 32 24 32; do
JJ_LEAF_SIZE=$leaf JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/it.json 2> gpurun_out/it.err
python -c "
import json
d=json.load(open('gpurun_out/it.json')); print('leaf $leaf us/timestep %.1f  %.2f Gjs/s'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9))" || tail -3 gpurun_out/it.err
done
