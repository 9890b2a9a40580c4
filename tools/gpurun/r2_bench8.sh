mkdir -p gpurun_out
N=${1:-8}
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -5 gpurun_out/r2_bench_${N}gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print('N=%d cfg2 us/timestep %.2f  %.2f Gjs/s frac %.3f e2e %.2f G'%(d['n_gpus'], d['ms_per_step']*1e3/d['config']['time_steps_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9))
for k,v in d['per_config'].items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','e2e','roofline_frac','device_us_per_time_step','setup_s','error','problems_this_gpu')})
print(d.get('shard_check'))
"
