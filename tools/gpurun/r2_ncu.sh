mkdir -p gpurun_out
# launch list of the bench command (kernel share of the step)
JJ_BENCH_INNER=100 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_launches.log 2>&1
tail -3 gpurun_out/r02_launches.log | cut -c1-300
# full capture of the step kernel: cfg2 (lean kernel), 10 time steps in the launch
JJ_BENCH_INNER=10 JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subdomain -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg2 python bench.py --steps 1 --warmup 1 > gpurun_out/r02_ncu_cfg2.log 2>&1
tail -2 gpurun_out/r02_ncu_cfg2.log | cut -c1-200
# full capture of the general kernel on cfg4 (512 problems, 4 time steps in the launch)
cat > /tmp/cfg4_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, pyjjasim_b200 as pj, bench
a, kw, Nt, (w0, w1), W_total, note = bench.named_config(pj, "cfg4", 0, 1)
for _ in range(2):
    pj.TimeEvolutionProblem(a, time_step_count=4, store_time_steps=[3], store_current=False, store_voltage=False, **kw).compute()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_subdomain -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg4 python /tmp/cfg4_run.py > gpurun_out/r02_ncu_cfg4.log 2>&1
tail -2 gpurun_out/r02_ncu_cfg4.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
