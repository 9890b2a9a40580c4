mkdir -p gpurun_out
for rep in 1 2; do
for lib in libjjstep.so libjjstep_a.so; do
JJ_LIB_PATH=$PWD/pyjjasim_b200/$lib JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ab.json 2> gpurun_out/r2_ab.err
python -c "
import json
d=json.load(open('gpurun_out/r2_ab.json')); print('$lib cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_ab.err
done
done
JJ_LIB_PATH=$PWD/pyjjasim_b200/libjjstep_a.so JJ_SUB_PROF=1 JJ_BENCH_INNER=300 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ab_prof.err
grep -A 12 "JJ_SUB_PROF" gpurun_out/r2_ab_prof.err | grep -v "sweep level" | tail -10 | cut -c1-160
