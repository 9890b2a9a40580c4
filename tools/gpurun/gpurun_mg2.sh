mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; tail -c 1500 gpurun_out/bench_g2.json; tail -3 gpurun_out/bench_g2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -c 600
timeout 600 python -m pytest tests -m gpu -x -q -k "shard or devices or distributed" 2>&1 | tail -2
