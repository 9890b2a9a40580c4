mkdir -p gpurun_out
run() {
  env "$@" JJ_BENCH_SKIP_E2E=1 JJ_BENCH_CONFIGS=cfg1,cfg3,cfg4 timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_knob.json 2> gpurun_out/r2_knob.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_knob.json')); pc=d['per_config']
print('$*'.ljust(24), 'cfg2 %.2f us %.3f | cfg1 %.1f us | cfg3 %.1f us %.3f | cfg4 %.1f us %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['roofline']['frac'], pc['cfg1']['device_us_per_time_step'], pc['cfg3']['device_us_per_time_step'], pc['cfg3']['roofline_frac'], pc['cfg4']['device_us_per_time_step'], pc['cfg4']['roofline_frac']))" || tail -3 gpurun_out/r2_knob.err
}
run JJ_LEAF_SIZE=16
run JJ_LEAF_SIZE=32
run JJ_LEAF_SIZE=16
run JJ_LEAF_SIZE=14
run JJ_LEAF_SIZE=15
run JJ_LEAF_SIZE=17
