mkdir -p gpurun_out
run() {
  env "$@" JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_knob.json 2> gpurun_out/r2_knob.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_knob.json')); print('$*'.ljust(44), 'cfg2 %.2f us  frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['roofline']['frac']))" || tail -3 gpurun_out/r2_knob.err
}
run X=1
run JJ_BAL_W=41,47,29
run JJ_BAL_W=60,45,28
run JJ_BAL_W=30,50,30
run JJ_BAL_ROUNDS=4
run JJ_BAL_ROUNDS=6 JJ_BAL_W=41,47,29
run JJ_BAL_ROUNDS=0
run JJ_LEAF_SIZE=17 JJ_BAL_W=41,47,29
