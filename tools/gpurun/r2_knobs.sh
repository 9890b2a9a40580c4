mkdir -p gpurun_out
run() {
  env "$@" JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_knob.json 2> gpurun_out/r2_knob.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_knob.json')); print('$*'.ljust(44), 'cfg2 %.2f us  frac %.3f  subdomains %s'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['roofline']['frac'], d['config']['subdomains_or_cluster']))" || tail -3 gpurun_out/r2_knob.err
}
run X=0
run JJ_LEAF_SIZE=24
run JJ_LEAF_SIZE=28
run JJ_LEAF_SIZE=36
run JJ_BENCH_NPARTS=17
run JJ_BENCH_NPARTS=16
run JJ_SUB_BALANCE=0
run JJ_SUB_NO_L2_WINDOW=1
run JJ_SUB_NO_TSLOT4=1
