mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "full_size or distribution or reference_style" > gpurun_out/r2_tests_new.txt 2>&1
tail -15 gpurun_out/r2_tests_new.txt
for lib in libjjstep.so libjjstep_f64.so; do
JJ_LIB_PATH=$PWD/pyjjasim_b200/$lib JJ_BENCH_SKIP_E2E=1 JJ_BENCH_SKIP_CONFIGS=1 timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_noise_$lib.json 2> gpurun_out/r2_noise_$lib.err
python -c "
import json
d=json.load(open('gpurun_out/r2_noise_$lib.json')); print('$lib cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f'%(d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac']))" || tail -5 gpurun_out/r2_noise_$lib.err
done
for c in cfg4 cfg3; do
timeout 600 python tools/e2e_profile_cfg.py $c 200 > gpurun_out/r2_e2eprof_$c.txt 2>&1
head -60 gpurun_out/r2_e2eprof_$c.txt | cut -c1-180
done
for tool in memcheck synccheck; do
for w in cfg2 upper; do
timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py $w > gpurun_out/r2_sanitizer_${tool}_$w.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_${tool}_$w.txt
done
done
