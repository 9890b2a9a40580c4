mkdir -p gpurun_out
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -5 gpurun_out/r2_bench_2gpu.err
cut -c1-4000 gpurun_out/r2_bench_2gpu.json
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "shards_over_devices" 2>&1 | tail -3
