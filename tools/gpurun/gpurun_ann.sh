set -x
mkdir -p gpurun_out
timeout 900 python tools/anneal_bench.py > gpurun_out/anneal_bench.json 2> gpurun_out/anneal_bench.err; tail -c 2000 gpurun_out/anneal_bench.json; tail -5 gpurun_out/anneal_bench.err
