mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_g8.json 2> gpurun_out/bench_g8.err; tail -c 1200 gpurun_out/bench_g8.json; tail -3 gpurun_out/bench_g8.err
