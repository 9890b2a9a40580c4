mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
ANNEAL_CPU=$((i-1)) timeout 900 python tools/anneal_bench.py 2> gpurun_out/anneal_bench.err > gpurun_out/anneal_bench.json; python -c "
import json
d=json.load(open('gpurun_out/anneal_bench.json')); print('device wall %.3f s  device_ms %.1f  %.2f G js/s'%(d['device']['wall_s'], d['device']['device_ms'], d['device']['junction_steps_per_s']/1e9), d.get('cpu',{}).get('junction_steps_per_s'))"
done
