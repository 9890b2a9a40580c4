mkdir -p gpurun_out
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
tail -5 gpurun_out/r2_bench_full.err
cut -c1-3000 gpurun_out/r2_bench_full.json
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -4 gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_ref.json
