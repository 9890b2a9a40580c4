mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/e2e_profile.py 2>&1 | head -18 | tail -14 | cut -c1-150
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; python -c "
import json
d=json.load(open('gpurun_out/bench_r01.json')); print('value %.2f G  e2e %.2f G  %.4f s/step'%(d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['seconds_per_step']))"
JJ_PINNED_RESULTS=0 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_np.json 2> gpurun_out/bench_np.err; python -c "
import json
d=json.load(open('gpurun_out/bench_np.json')); print('unpinned: value %.2f G  e2e %.2f G  %.4f s/step'%(d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['seconds_per_step']))"
