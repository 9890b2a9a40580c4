mkdir -p gpurun_out
for rep in 1 2; do
for lib in libjjstep.so libjjstep_r2.so libjjstep_r0.so; do
JJ_LIB_PATH=$PWD/pyjjasim_b200/$lib JJ_BENCH_SKIP_E2E=1 JJ_BENCH_CONFIGS=cfg3,cfg4 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_ab2.json 2> gpurun_out/r2_ab2.err
python -c "
import json
d=json.load(open('gpurun_out/r2_ab2.json')); pc=d['per_config']
print('$lib cfg2 %.2f us frac %.3f | cfg3 %.1f us %.3f | cfg4 %.1f us %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['roofline']['frac'], pc['cfg3']['device_us_per_time_step'], pc['cfg3']['roofline_frac'], pc['cfg4']['device_us_per_time_step'], pc['cfg4']['roofline_frac']))" || tail -5 gpurun_out/r2_ab2.err
done
done
